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


@pytest.mark.gpu
def test_realign_bam_end_to_end(golden, tables, tmp_path):
    """BAM + FASTA in, realigned SAM out (header + records in coordinate order) == the reference's golden records."""
    bam, fasta, reads, g = _fixture(golden, tmp_path)
    cfg.args.sub_scores, cfg.args.np_scores = tables
    cfg.args.max_n, cfg.args.max_l = 6, 100
    lines = bamio.realign_bam(bam, fasta, out_prefix=str(tmp_path / "out"), argv=["realign.py"])
    want = {l.split("\t")[0]: l for l in g["expected_sam"]}
    assert lines == [want[r[0]] for r in reads]
    body = [l for l in open(str(tmp_path / "out.sam")).read().splitlines() if not l.startswith("@")]
    assert body == lines
