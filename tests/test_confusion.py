"""Score-table construction and basecaller confusion matrices (A14 "produced" side and N4).

CPU tests: aln.calc_score_matrices / fix_matrix_properties against the reference's own outputs (bit-exact); the CPU oracle
of the pileup path (oracle/pileup_oracle.py) against vectors produced by the compiled reference's
bam.calc_confusion_matrices (bam.pyx:351-510); host packing.  GPU tests (-m gpu): npore_confusion_batch through the C ABI
against the same golden vectors and against the oracle on seeded inputs.  Counts are integers: everything is exact."""
import os

import numpy as np
import pytest

import oracle
import pileup_oracle as po
from npore_b200 import aln, cfg, confusion, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_calc_score_matrices_matches_reference_tables(tables):
    """guppy5 counts -> tables: identical to the bit with what the reference built from the same counts (tables.npz)."""
    cm = np.load(os.path.join(GOLD, "cm_guppy5.npz"))
    sub, npt, ins, dele = aln.calc_score_matrices(cm["subs"], cm["nps"], cm["inss"], cm["dels"])
    assert np.array_equal(_bits(sub), _bits(tables[0])) and np.array_equal(_bits(npt), _bits(tables[1]))
    assert ins.dtype == np.float32 and ins.shape == cm["inss"].shape and ins[-1] == 0 and dele[-1] == 0
    assert np.all(np.diagonal(npt, axis1=1, axis2=2)[:, 1:] == 0) and np.all(sub[0] == 0) and np.all(sub[:, 0] == 0)


def test_calc_score_matrices_live_against_reference():
    import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not built here")
    ref = ref_loader.load_reference()
    rng = np.random.default_rng(5)
    for trial in range(3):
        nps = rng.integers(0, 50, size=(6, 101, 101)) * (rng.random((6, 101, 101)) < 0.2)
        nps[:, np.arange(101), np.arange(101)] += rng.integers(0, 5000, size=(6, 101))
        if trial == 2:
            nps[:, 40:] = 0                      # rows without observations
        subs = rng.integers(0, 10000, size=(5, 5)); inss = rng.integers(0, 1000, size=101); dels = rng.integers(0, 1000, size=101)
        for a, b in zip(aln.calc_score_matrices(subs, nps, inss, dels), ref.aln.calc_score_matrices(subs, nps, inss, dels)):
            assert np.array_equal(_bits(a), _bits(b))


def _case_reads(c):
    return [(r[0], [(n, op) for n, op in r[1]], r[2], None if r[3] is None else bytes(ord(ch) - 33 for ch in r[3]), r[4], r[5])
            for r in c["reads"]]


def _case_want(c, max_n=6, max_l=100):
    nps = np.zeros((max_n, max_l + 1, max_l + 1), np.int64)
    for a, b, d, v in c["nps"]:
        nps[a, b, d] = v
    return np.array(c["subs"], np.int64), nps, np.array(c["inss"], np.int64), np.array(c["dels"], np.int64)


def test_pileup_oracle_matches_reference_vectors(golden):
    """oracle/pileup_oracle.py end to end (restated mpileup text -> restated parser) == the compiled reference's parser
    on the same text (tests/golden/confusion_kats.json.gz).  Pins the parser; the mpileup restatement itself is unpinned."""
    cases = golden("confusion_kats.json.gz")
    events = 0
    for c in cases:
        got = po.confusion([po.Read(*r) for r in _case_reads(c)], c["contig"], c["start"], c["end"], oracle.get_np_info, oracle.bases_to_int)
        for a, b in zip(_case_want(c), got):
            assert np.array_equal(a, b)
        events += int(got[1].sum() - np.trace(got[1], axis1=1, axis2=2).sum())
    assert events > 300          # copy-number events actually occur in the fixtures


def test_pileup_text_examples():
    """Hand-checked lines of the mpileup restatement (samtools conventions: '^' + mapq char, '$', '*', '-2NN' / '+2AG' on
    the position before the op, '*+' after a deletion, -Q filter on the entry's query position)."""
    R = po.Read
    hi = bytes([40] * 12)
    r1 = R(2, [(3, "M"), (2, "D"), (2, "M"), (2, "I"), (1, "M")], "ACGTAGGC", hi[:8], 0, 60)
    r2 = R(3, [(2, "S"), (2, "M"), (1, "D"), (1, "I"), (2, "M")], "TTCGAAC", None, 16, 0)
    lines = po.mpileup_column5([r1, r2], 0, 12)
    assert lines == ["^]A", "C^!C", "G-2NNG-1N", "**+1A", "*A", "TC$", "A+2GG", "C$"]
    low = bytes([40, 40, 5, 40, 40, 40, 40, 40])
    assert po.mpileup_column5([R(2, r1.cigar, r1.seq, low, 0, 60)], 0, 12)[2] == "*"           # the only entry is filtered
    assert po.mpileup_column5([R(0, [(2, "M")], "AC", None, 0x400, 60)], 0, 5) == []           # duplicate: dropped


def test_host_packing():
    ar = confusion.AlignedReads([(50, [(10, "M")], "A" * 10, None, 0), (5, [(3, "S"), (20, "M"), (5, "D"), (4, "M")], "C" * 27, bytes(27), 0),
                                 (30, [(10, "=")], "G" * 10, None, 0x100), (20, [(100, "M")], "T" * 100, None, 16)])
    assert ar.pos.tolist() == [5, 20, 50] and ar.end.tolist() == [34, 120, 60] and ar.maxend.tolist() == [34, 120, 120]
    assert ar.overlapping(0, 5).tolist() == [] and ar.overlapping(0, 6).tolist() == [0] and ar.overlapping(34, 50).tolist() == [1]
    assert ar.overlapping(119, 500).tolist() == [1] and ar.overlapping(55, 56).tolist() == [1, 2]
    assert ar.qual[:27].tolist() == [0] * 27 and ar.qual[27] == 255
    old = cfg.args.chunk_width
    try:
        cfg.args.chunk_width = 100
        assert confusion.get_ranges([("a", 0, 250), ("b", 10, 20)]) == [("a", 0, 100), ("a", 100, 200), ("a", 200, 250), ("b", 10, 20)]
    finally:
        cfg.args.chunk_width = old
    pk = confusion.PileupPack([("a", 0, 40), ("a", 40, 130)], {"a": "ACGT" * 40}, {"a": ar}, 6)
    assert pk.ref_off.tolist() == [0, 47, 47 + 97] and pk.range_reads.tolist() == [0, 1, 1, 2] and pk.range_reads_off.tolist() == [0, 2, 4]
    with pytest.raises(ValueError):
        confusion.PileupPack([("a", 0, 161)], {"a": "ACGT" * 40}, {"a": ar}, 6)


# ------------------------------------------------------------------------------------------------ GPU
def _gpu_counts(contig, reads, ranges):
    return confusion.calc_confusion_matrices_batch(ranges, refs={"c": contig}, reads={"c": confusion.AlignedReads([r[:5] for r in reads])})


@pytest.mark.gpu
def test_gpu_confusion_golden(golden):
    """npore_confusion_batch == the compiled reference's calc_confusion_matrices on every golden window."""
    for c in golden("confusion_kats.json.gz"):
        got = _gpu_counts(c["contig"], _case_reads(c), [("c", c["start"], c["end"])])
        for nm, a, b in zip(("subs", "nps", "inss", "dels"), _case_want(c), got):
            assert np.array_equal(a, b), f"seed {c['seed']}: {nm} differs"


@pytest.mark.gpu
def test_gpu_confusion_fuzz_vs_oracle():
    """Seeded windows incl. lower-case reference, uncovered stretches, >32-deep pileups, several windows per call
    (the sum over windows is what get_confusion_matrices reduces to, bam.pyx:186-192), contig-end clipping."""
    for seed in range(40):
        rng = np.random.default_rng(9100 + seed)
        L = int(rng.integers(300, 3000))
        contig = synth.make_reference(L, rng, p_np=0.4)
        if seed % 4 == 1:
            contig = contig[:L // 2] + contig[L // 2:].lower()
        reads = synth.make_aligned_reads(contig, int(rng.integers(1, 60)) if seed % 5 else 250, int(rng.integers(50, 900)), rng)
        start, end = (0, L) if seed % 3 == 0 else sorted(int(x) for x in rng.choice(L + 1, size=2, replace=False))
        cuts = sorted({start, end, *(int(x) for x in rng.integers(start, end + 1, size=int(rng.integers(0, 4))))})
        ranges = [("c", a, b) for a, b in zip(cuts[:-1], cuts[1:])]
        want = None
        for _, a, b in ranges:
            one = po.confusion([po.Read(*r) for r in reads], contig, a, b, oracle.get_np_info, oracle.bases_to_int)
            want = one if want is None else tuple(x + y for x, y in zip(want, one))
        got = _gpu_counts(contig, reads, ranges)
        for nm, a, b in zip(("subs", "nps", "inss", "dels"), want, got):
            assert np.array_equal(a, b), f"seed {seed}: {nm} differs"


@pytest.mark.gpu
def test_gpu_confusion_from_bam_and_scale(tmp_path, tables):
    """End to end like the reference's `--recalc_cms` run (bam.pyx:176-199): BAM file -> get_ranges windows -> counts -> cached
    .npy -> calc_score_matrices; at a size (2 Mb of aligned bases, 40 windows) where only invariants are checked on the
    whole and the oracle on a sample window."""
    from npore_b200 import bamio
    rng = np.random.default_rng(77)
    contig = synth.make_reference(60000, rng, p_np=0.35)
    reads = synth.make_aligned_reads(contig, 700, 3000, rng, with_clips=False)
    recs = [{"name": f"r{k}", "flag": r[4], "ref_id": 0, "pos": r[0], "mapq": r[5], "cigar": r[1], "seq": r[2], "qual": r[3]}
            for k, r in enumerate(reads)]
    bam_fn = str(tmp_path / "cm.bam")
    bamio.write_bam(bam_fn, "@HD\tVN:1.6\tSO:coordinate\n", [("c", len(contig))], recs)
    saved = {k: getattr(cfg.args, k, None) for k in ("bam", "refs", "regions", "chunk_width", "stats_dir", "recalc_cms", "_alignments")}
    try:
        cfg.args.bam, cfg.args.refs, cfg.args.regions = bam_fn, {"c": contig}, [("c", 0, len(contig))]
        cfg.args.chunk_width, cfg.args.stats_dir, cfg.args.recalc_cms, cfg.args._alignments = 1500, str(tmp_path / "stats"), True, None
        subs, nps, inss, dels = confusion.get_confusion_matrices()
        assert os.path.exists(tmp_path / "stats" / "nps_cm.npy")
        one = confusion.calc_confusion_matrices(("c", 3000, 4500))
        want = po.confusion([po.Read(*r) for r in reads], contig, 3000, 4500, oracle.get_np_info, oracle.bases_to_int)
        for a, b in zip(want, one):
            assert np.array_equal(a, b)
        # every counted base entry closes exactly one window: inss[0] + (windows with an insertion) == number of base entries
        n_base = int(subs.sum())
        assert 0 < inss[0] <= n_base and 0 < dels[0] <= n_base and subs[1:, 1:].trace() > 0.9 * n_base
        assert n_base - inss[0] <= inss[1:].sum() + (nps.sum() - np.trace(nps, axis1=1, axis2=2).sum())
        cfg.args.recalc_cms = False
        again = confusion.get_confusion_matrices()
        assert all(np.array_equal(a, b) for a, b in zip(again, (subs, nps, inss, dels)))
        sub_scores, np_scores, _, _ = aln.calc_score_matrices(subs, nps, inss, dels)
        assert sub_scores.shape == (5, 5) and np_scores.shape == nps.shape and np.isfinite(np_scores).all()
    finally:
        for k, v in saved.items():
            setattr(cfg.args, k, v)


@pytest.mark.gpu
def test_gpu_confusion_error_behaviour():
    ar = confusion.AlignedReads([(0, [(4, "M"), (2, "P"), (4, "M")], "ACGTACGT", None, 0)])
    with pytest.raises(RuntimeError):
        confusion.calc_confusion_matrices_batch([("c", 0, 8)], refs={"c": "ACGTACGTAC"}, reads={"c": ar})
    empty = confusion.calc_confusion_matrices_batch([("c", 0, 8)], refs={"c": "ACGTACGTAC"}, reads={})
    assert all(int(m.sum()) == 0 for m in empty)
