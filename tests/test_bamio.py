"""N1/N2: BAM ingest without pysam and the ordered SAM writer, the callers either side of the hot path."""
import os
import re

import numpy as np
import pytest

from npore_b200 import bamio, cfg


def _fixture(golden, tmp_path):
    """A coordinate-sorted BAM + FASTA equivalent to the reference's test/data (rebuilt from the golden read tuples)."""
    g = golden("golden_sam.json")
    reads = sorted(g["reads"], key=lambda r: r[3])
    contig = ["N"] * 1001
    for r in reads:
        contig[r[3]:r[6]] = r[9]
    fasta = tmp_path / "ref.fasta"
    fasta.write_text(">ref test contig\n" + "\n".join("".join(contig)[i:i + 60] for i in range(0, 1001, 60)) + "\n")
    recs = []
    for r in reads:
        name, flag, _, start, mapq, cigar, _, seq, quals, _, hap = r
        recs.append({"name": name, "flag": flag, "ref_id": 0, "pos": start, "mapq": mapq,
                     "cigar": [(int(n), op) for n, op in re.findall(r"(\d+)(\D)", cigar)], "seq": seq,
                     "qual": None if quals == "*" else bytes(ord(c) - 33 for c in quals), "tags": {"HP": hap} if hap else {}})
    # plus records the ingest must skip (bam.pyx:31-32) and one with soft clips
    recs.append({"name": "secondary", "flag": 256, "ref_id": 0, "pos": 5, "mapq": 0, "cigar": [(4, "=")], "seq": "ACGT", "qual": None})
    recs.append({"name": "unmapped", "flag": 4, "ref_id": -1, "pos": -1, "mapq": 0, "cigar": [], "seq": "ACGT", "qual": None})
    bam = tmp_path / "reads.bam"
    bamio.write_bam(str(bam), "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:ref\tLN:1001\n", [("ref", 1001)], recs)
    return str(bam), str(fasta), reads, g


def test_bam_roundtrip_and_filters(golden, tmp_path):
    bam, fasta, reads, _ = _fixture(golden, tmp_path)
    text, refs, recs = bamio.read_bam(bam)
    assert refs == [("ref", 1001)] and text.startswith("@HD")
    assert len(list(recs)) == len(reads) + 2
    got = list(bamio.get_read_data(bam, fasta))
    assert [list(t) for t in got] == reads                        # secondary / unmapped dropped, tuples as bam.pyx:34-47
    assert len(list(bamio.get_read_data(bam, fasta, max_reads=3))) == 3
    sub = list(bamio.get_read_data(bam, fasta, regions=[("ref", 0, reads[0][6])]))
    assert sub and all(t[3] < reads[0][6] for t in sub)


def test_soft_clips_are_stripped(tmp_path):
    fasta = tmp_path / "r.fa"
    fasta.write_text(">c\nACGTACGTACGT\n")
    bam = tmp_path / "s.bam"
    bamio.write_bam(str(bam), "@HD\tVN:1.6\n", [("c", 12)], [
        {"name": "x", "flag": 0, "ref_id": 0, "pos": 2, "mapq": 9, "cigar": [(2, "S"), (4, "="), (1, "S")], "seq": "ttGTACg",
         "qual": bytes([1, 2, 3, 4, 5, 6, 7]), "tags": {"HP": 2}}])
    (t,) = list(bamio.get_read_data(str(bam), str(fasta)))
    assert t == ("x", 0, "c", 2, 9, "2S4=1S", 6, "GTAC", "".join(chr(33 + q) for q in (3, 4, 5, 6)), "GTAC", 2)


def test_native_reader_matches_python_reader(golden, tmp_path):
    """libnpore_b200.so's BAM ingest (include/npore_bamio.h; host code, no GPU needed) == the pure-Python decode, incl. soft /
    hard clips, missing qualities, lower-case and IUPAC bases, HP tags of several integer types, filtered records."""
    from npore_b200 import cig
    bam, fasta, reads, _ = _fixture(golden, tmp_path)
    extra = tmp_path / "x.bam"
    recs = [{"name": "clip", "flag": 16, "ref_id": 0, "pos": 2, "mapq": 9, "cigar": [(3, "H"), (2, "S"), (4, "="), (1, "I"), (2, "D"), (3, "X"), (1, "S")],
             "seq": "ttGTACRacng", "qual": bytes(range(1, 12)), "tags": {"HP": 2}},
            {"name": "noq", "flag": 0, "ref_id": 0, "pos": 7, "mapq": 60, "cigar": [(5, "M")], "seq": "ACGTN", "qual": None},
            {"name": "sec", "flag": 0x100, "ref_id": 0, "pos": 8, "mapq": 1, "cigar": [(2, "M")], "seq": "AC", "qual": None},
            {"name": "other", "flag": 0, "ref_id": 1, "pos": 0, "mapq": 3, "cigar": [(4, "S"), (2, "M")], "seq": "ACGTAC", "qual": bytes(6)}]
    bamio.write_bam(str(extra), "@HD\tVN:1.6\n", [("c", 30), ("d", 9)], recs)
    fa2 = {"c": "ACGTACGTACGTACGTACGTACGTACGTAC", "d": "ACGTACGTA"}
    for path, fa in ((bam, bamio.read_fasta(fasta)), (str(extra), fa2)):
        nb = bamio.NativeBam(path, n_threads=3)
        text, refs, _ = bamio.read_bam(path)
        assert nb.refs == refs and nb.text == text
        want = list(bamio.get_read_data(path, fa))
        sels = list(bamio.select_reads(nb))
        sel = np.concatenate([s for _, s in sels])
        assert len(sel) == len(want)
        g = nb.gather(sel, n_threads=2)
        for k, w in enumerate(want):
            cut = lambda a, o: a[g[o][k]:g[o][k + 1]]   # noqa: E731
            assert cut(g["names"], "name_off").tobytes().decode() == w[0]
            assert (nb.flag[sel[k]], nb.refs[nb.ref_id[sel[k]]][0], nb.pos[sel[k]], nb.mapq[sel[k]], nb.end[sel[k]], nb.hp[sel[k]]) == (w[1], w[2], w[3], w[4], w[6], w[10])
            assert cut(g["seq_ascii"], "seq_off").tobytes().decode() == w[7]
            assert np.array_equal(cut(g["seq_codes"], "seq_off"), cig.bases_to_int(w[7]))
            assert (cut(g["qual_ascii"], "seq_off").tobytes().decode() if nb.has_qual[sel[k]] else "*") == w[8]
            assert "".join(f"{int(x) >> 4}{'MIDNSHP=XB'[int(x) & 15]}" for x in cut(g["cigar"], "cig_off")) == re.sub(r"\d+[SH]", "", w[5])
        # SAM text of the batch formatter == bam.sam_record, with the (clip-free) input CIGARs standing in for results
        from npore_b200.bam import sam_record
        from npore_b200.engine import cigars_to_rle_batch
        kept = [re.sub(r"\d+[SH]", "", w[5]) for w in want]
        words, off = cigars_to_rle_batch(kept)
        blob = bamio.format_sam(nb, sel, g, words, off, n_threads=2).tobytes().decode()
        assert blob == "".join(sam_record(w, c) + "\n" for w, c in zip(want, kept))
        # npore_sam_format_fd: the same block, appended to a file by the formatter threads themselves
        outp = str(tmp_path / "fd.sam")
        with open(outp, "wb") as fh:
            fh.write(b"@HD\n")
        fd = os.open(outp, os.O_WRONLY)
        got = bamio.format_sam(nb, sel, g, words, off, n_threads=3, fd=fd, offset=4)
        os.close(fd)
        assert open(outp, "rb").read() == b"@HD\n" + blob.encode() and len(got) == len(blob)
        assert len(list(bamio.select_reads(nb, max_reads=1))[0][1]) == 1
        # streaming: small windows (records straddle BGZF members and window ends) deliver the same records in order
        st = bamio.NativeBam(path, n_threads=2, window_bytes=120)
        seen, windows = [], 0
        while st.advance():
            windows += 1
            gg = st.gather(np.arange(st.n))
            seen += [(int(st.pos[k]), int(st.flag[k]), gg["seq_ascii"][gg["seq_off"][k]:gg["seq_off"][k + 1]].tobytes()) for k in range(st.n)]
        allg = nb.gather(np.arange(nb.n))
        assert seen == [(int(nb.pos[k]), int(nb.flag[k]), allg["seq_ascii"][allg["seq_off"][k]:allg["seq_off"][k + 1]].tobytes()) for k in range(nb.n)]
        assert windows > 1 or nb.n < 3
        st.close()
        nb.close()
    with pytest.raises(FileNotFoundError):
        bamio.NativeBam(str(tmp_path / "missing.bam"))
    bad = tmp_path / "bad.bam"
    bad.write_bytes(open(bam, "rb").read()[:200])
    with pytest.raises(ValueError):
        bamio.NativeBam(str(bad))


def test_native_reader_random_stress(tmp_path):
    """Seeded random BAMs (empty reads, clips, all op kinds, missing qualities, long names, filtered flags) streamed in
    random window sizes on random thread counts: every record equals the pure-Python decode."""
    rng = np.random.default_rng(5)
    for trial in range(25):
        recs = []
        for k in range(int(rng.integers(0, 40))):
            L = int(rng.integers(0, 400))
            seq = "".join(rng.choice(list("ACGTNacgtRY"), size=L)) if L else ""
            ops, rem, tail = [], L, 0
            if L and rng.random() < 0.3:
                ops.append((int(rng.integers(1, 5)), "H"))
            if rem > 4 and rng.random() < 0.4:
                c = int(rng.integers(1, 4)); ops.append((c, "S")); rem -= c
            if rem > 4 and rng.random() < 0.4:
                tail = int(rng.integers(1, 4)); rem -= tail
            while rem > 0:
                m = int(rng.integers(1, rem + 1)); ops.append((m, "M=X"[int(rng.integers(0, 3))])); rem -= m
                if rem > 0 and rng.random() < 0.5:
                    ops.append((int(rng.integers(1, 4)), "D"))
                if rem > 1 and rng.random() < 0.5:
                    i = int(rng.integers(1, min(3, rem))); ops.append((i, "I")); rem -= i
            if tail:
                ops.append((tail, "S"))
            merged = []
            for ln, op in ops:
                if merged and merged[-1][1] == op:
                    merged[-1] = (merged[-1][0] + ln, op)
                else:
                    merged.append((ln, op))
            recs.append({"name": f"r{k}_{'x' * int(rng.integers(0, 20))}", "flag": int(rng.choice([0, 16, 256, 4, 2048, 1024])),
                         "ref_id": int(rng.integers(0, 2)), "pos": int(rng.integers(0, 900)), "mapq": int(rng.integers(0, 61)), "cigar": merged,
                         "seq": seq, "qual": None if (rng.random() < 0.3 or L == 0) else bytes(rng.integers(0, 60, size=L).astype(np.uint8)),
                         "tags": {"HP": int(rng.integers(0, 3))} if rng.random() < 0.5 else {}})
        path = str(tmp_path / f"s{trial}.bam")
        bamio.write_bam(path, "@HD\tVN:1.6\n", [("a", 5000), ("b", 4000)], recs)
        py = list(bamio.read_bam(path)[2])
        nb = bamio.NativeBam(path, n_threads=int(rng.integers(1, 5)), window_bytes=int(rng.integers(40, 3000)))
        k = 0
        while nb.advance():
            g = nb.gather(np.arange(nb.n))
            for i in range(nb.n):
                r = py[k]; k += 1
                ops, lens = r["cigar"] & 15, r["cigar"] >> 4
                lead = int(lens[0]) if len(ops) and ops[0] == 4 else (int(lens[1]) if len(ops) > 1 and ops[0] == 5 and ops[1] == 4 else 0)
                trail = int(lens[-1]) if len(ops) and ops[-1] == 4 else (int(lens[-2]) if len(ops) > 1 and ops[-1] == 5 and ops[-2] == 4 else 0)
                cut = lambda a, o: a[g[o][i]:g[o][i + 1]]   # noqa: E731
                assert cut(g["seq_ascii"], "seq_off").tobytes().decode() == r["seq"][lead:len(r["seq"]) - trail].upper()
                assert cut(g["names"], "name_off").tobytes().decode() == r["name"]
                assert (nb.pos[i], nb.flag[i], nb.mapq[i], nb.hp[i]) == (r["pos"], r["flag"], r["mapq"], int(r["tags"].get("HP", 0)))
                assert nb.end[i] == r["pos"] + int(lens[np.isin(ops, (0, 2, 3, 7, 8))].sum())
                assert np.array_equal(cut(g["cigar"], "cig_off"), r["cigar"][(ops != 4) & (ops != 5)])
                if r["qual"] is not None and len(r["seq"]):
                    assert nb.has_qual[i] == 1 and cut(g["qual_ascii"], "seq_off").tobytes() == (r["qual"][lead:len(r["qual"]) - trail] + np.uint8(33)).tobytes()
                elif len(r["seq"]):
                    assert nb.has_qual[i] == 0
        assert k == len(py) and nb.advance() == 0
        nb.close()


def test_header(tmp_path):
    out = tmp_path / "d" / "o.sam"
    bamio.create_header(str(out), [("chr1", 100), ("chr2", 50)], argv=["realign.py", "--bam", "x"])
    lines = out.read_text().splitlines()
    assert lines[0] == "@HD\tVN:1.6\tSO:coordinate" and lines[1:3] == ["@SQ\tSN:chr1\tLN:100", "@SQ\tSN:chr2\tLN:50"]
    assert lines[3].startswith("@PG\tID:realigner\tPN:realigner") and lines[3].endswith("CL:realign.py --bam x")


def test_reference_bam_fixture_live(golden):
    """Where /root/reference exists: pure-Python decode of test/data/reads.bam == the records of reads.sam."""
    bam, fa = "/root/reference/test/data/reads.bam", "/root/reference/test/data/ref.fasta"
    if not os.path.exists(bam):
        pytest.skip("reference tree not present")
    by = {r[0]: r for r in golden("golden_sam.json")["reads"]}
    got = list(bamio.get_read_data(bam, fa))
    assert len(got) == 10 and all(list(t) == by[t[0]] for t in got)
    nb = bamio.NativeBam(bam)                               # the native reader on the reference's own file
    sel = np.concatenate([s for _, s in bamio.select_reads(nb)])
    g = nb.gather(sel)
    for k, t in enumerate(got):
        assert g["seq_ascii"][g["seq_off"][k]:g["seq_off"][k + 1]].tobytes().decode() == t[7]
        assert (g["qual_ascii"][g["seq_off"][k]:g["seq_off"][k + 1]].tobytes().decode() if nb.has_qual[sel[k]] else "*") == t[8]
        assert (int(nb.pos[sel[k]]), int(nb.end[sel[k]]), int(nb.hp[sel[k]]), int(nb.flag[sel[k]])) == (t[3], t[6], t[10], t[1])
    nb.close()


@pytest.mark.gpu
def test_realign_bam_end_to_end(golden, tables, tmp_path):
    """BAM + FASTA in, realigned SAM out (header + records in coordinate order) == the reference's golden records."""
    bam, fasta, reads, g = _fixture(golden, tmp_path)
    cfg.args.sub_scores, cfg.args.np_scores = tables
    cfg.args.max_n, cfg.args.max_l = 6, 100
    n = bamio.realign_bam(bam, fasta, out_prefix=str(tmp_path / "out"), argv=["realign.py"])
    want = {l.split("\t")[0]: l for l in g["expected_sam"]}
    text = open(str(tmp_path / "out.sam")).read().splitlines()
    assert n == len(reads) and text[0] == "@HD\tVN:1.6\tSO:coordinate" and text[1] == "@SQ\tSN:ref\tLN:1001"
    assert [l for l in text if not l.startswith("@")] == [want[r[0]] for r in reads]
    # the tuple API (bam.pyx:18-89 call for call) writes the same records
    from npore_b200 import bam as nbam
    cfg.args.out_prefix = str(tmp_path / "out2")
    assert nbam.realign_reads(bamio.get_read_data(bam, fasta), write=False) == [want[r[0]] for r in reads]
    # several GPU batches and a region restriction
    n2 = bamio.realign_bam(bam, fasta, out_prefix=str(tmp_path / "out3"), argv=["x"], max_batch_ops=1500, regions=[("ref", 0, reads[3][6])])
    body3 = [l for l in open(str(tmp_path / "out3.sam")).read().splitlines() if not l.startswith("@")]
    assert n2 == len(body3) and body3 == [want[r[0]] for r in reads if r[3] < reads[3][6]]
    # streamed in small windows, three batches in flight, capped read count
    n3 = bamio.realign_bam(bam, fasta, out_prefix=str(tmp_path / "out4"), argv=["x"], max_batch_ops=1200, window_bytes=3000, n_inflight=3)
    body4 = [l for l in open(str(tmp_path / "out4.sam")).read().splitlines() if not l.startswith("@")]
    assert n3 == len(reads) and body4 == [want[r[0]] for r in reads]
    n4 = bamio.realign_bam(bam, fasta, out_prefix=str(tmp_path / "out5"), argv=["x"], window_bytes=3000, max_reads=4)
    assert n4 == 4 and [l for l in open(str(tmp_path / "out5.sam")).read().splitlines() if not l.startswith("@")] == [want[r[0]] for r in reads[:4]]


def test_gather_nib_is_the_file_packing(tmp_path):
    """npore_bam_gather_nib: the aligned bases as they lie in the record (4 bits each, soft clips skipped) -- decoding the nibbles
    with the table of include/npore_b200.h gives npore_bam_gather's seq_codes."""
    rng = np.random.default_rng(12)
    recs = []
    for k in range(60):
        n = int(rng.integers(0, 90))
        lead, trail = (int(rng.integers(0, 6)), int(rng.integers(0, 6))) if n > 12 else (0, 0)
        seq = "".join(rng.choice(list("ACGTNRY"), size=n))
        cig = ([(lead, "S")] if lead else []) + ([(n - lead - trail, "M")] if n - lead - trail else []) + ([(trail, "S")] if trail else [])
        recs.append({"name": f"r{k}", "flag": 0, "ref_id": 0, "pos": 10 * k, "mapq": 9, "cigar": cig, "seq": seq, "qual": None, "tags": {}})
    path = str(tmp_path / "n.bam")
    bamio.write_bam(path, "@HD\tVN:1.6\n", [("a", 5000)], recs)
    nb = bamio.NativeBam(path)
    sel = np.arange(nb.n)[::-1].copy()                    # any order
    g = nb.gather(sel)
    nib, start = nb.gather_nib(sel)
    lut = np.zeros(16, np.uint8); lut[[1, 2, 4, 8]] = [1, 2, 3, 4]
    for k in range(len(sel)):
        n = int(g["seq_off"][k + 1] - g["seq_off"][k])
        idx = start[k] + np.arange(n)
        b = nib[idx >> 1] if n else np.zeros(0, np.uint8)
        v = np.where(idx & 1, b & 15, b >> 4)
        assert np.array_equal(lut[v], g["seq_codes"][g["seq_off"][k]:g["seq_off"][k + 1]])
    # the text-only gather (two letters per packed byte; odd and even clip lengths, odd and even read lengths) == the per-base one
    gt = nb.gather(sel, want_codes=False, want_cigar=False)
    assert np.array_equal(gt["seq_ascii"][:int(gt["seq_off"][-1])], g["seq_ascii"][:int(g["seq_off"][-1])])
    nb.close()


def test_malformed_inputs_are_rejected_not_read_out_of_bounds(tmp_path):
    """a record with l_seq < 0 is treated as empty, a BGZF member claiming > 64 KiB is a format error, and
    npore_sam_format prints RNAME '*' for a ref_id outside the header's contigs."""
    import struct
    from npore_b200 import _lib
    recs = [{"name": "ok", "flag": 0, "ref_id": 0, "pos": 5, "mapq": 1, "cigar": [(4, "M")], "seq": "ACGT", "qual": None, "tags": {}}]
    path = str(tmp_path / "m.bam")
    bamio.write_bam(path, "@HD\tVN:1.6\n", [("a", 100)], recs)
    # rebuild the stream by hand with l_seq = -5 in the record
    body = bytearray()
    text = b"@HD\tVN:1.6\n"
    body += b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", 1) + struct.pack("<i", 2) + b"a\0" + struct.pack("<i", 100)
    rec = struct.pack("<iiBBHHHiiii", 0, 5, 3, 1, 0, 0, 0, -5, -1, -1, 0) + b"ok\0"
    body += struct.pack("<i", len(rec)) + rec
    with open(path, "wb") as fh:
        fh.write(bamio._bgzf_block(bytes(body))); fh.write(bamio._bgzf_block(b""))
    nb = bamio.NativeBam(path)
    assert nb.n == 1 and int(nb.aln_len[0]) == 0 and int(nb.n_cigar[0]) == 0
    g = nb.gather(np.arange(1))
    # ref_id out of range -> '*'
    cols = bamio.take_columns(nb, np.arange(1)); cols["ref_id"][:] = 7
    blob = bamio.format_sam(nb, None, g, np.zeros(1, np.uint32), np.zeros(2, np.int64), cols=cols).tobytes().decode()
    assert blob.split("\t")[2] == "*"
    nb.close()
    # ISIZE > 64 KiB
    blk = bytearray(bamio._bgzf_block(b"x" * 100))
    blk[-4:] = struct.pack("<I", 1 << 20)
    with open(path, "wb") as fh:
        fh.write(bytes(blk))
    with pytest.raises(ValueError):
        bamio.NativeBam(path)
    assert "npore_bam_gather_nib" in _lib.IO_EXPORTS


def test_short_last_window_is_merged_and_prefetch_keeps_the_order(tmp_path):
    """Streaming reader: the window after the current one is inflated in the background (npore_bam_prefetch) and a last window
    that would hold less than 0.4 windows of records is taken by its predecessor (a short last batch costs a full chunk latency
    on the GPU).  The records still arrive once each, in file order."""
    rng = np.random.default_rng(5)
    recs = [{"name": f"r{k:05d}", "flag": 0, "ref_id": 0, "pos": 3 * k, "mapq": 20, "cigar": [(150, "M")], "seq": "".join(rng.choice(list("ACGT"), size=150)),
             "qual": bytes([30] * 150), "tags": {"HP": k % 3}} for k in range(1000)]
    path = str(tmp_path / "w.bam")
    bamio.write_bam(path, "@HD\tVN:1.6\tSO:coordinate\n", [("a", 10_000)], recs)
    whole = bamio.NativeBam(path)
    per_record = 4 + 32 + 7 + 4 + 75 + 150 + 7                     # block_size field + fixed part + name + 1 cigar word + packed bases + quals + HP tag
    total = per_record * len(recs)
    window = int(total / 2.25)                                      # naive cut: two full windows and a quarter
    st = bamio.NativeBam(path, n_threads=2, window_bytes=window)
    sizes, names = [], []
    while st.advance():
        sizes.append(st.n)
        g = st.gather(np.arange(st.n), want_codes=False, want_cigar=False)
        names += [g["names"][g["name_off"][k]:g["name_off"][k + 1]].tobytes().decode() for k in range(st.n)]
    assert names == [r["name"] for r in recs] and sum(sizes) == whole.n == len(recs)
    assert len(sizes) == 2 and min(sizes) > 0.4 * window / per_record, sizes      # no quarter window at the end
    st.close(); whole.close()

