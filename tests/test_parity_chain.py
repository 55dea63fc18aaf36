"""GPU parity, closing the chain on the GPU box (-m gpu): the CUDA path through the C ABI against

  * the COMPILED, UNMODIFIED reference itself (oracle/_ref: aln.align with the 3-line score patch, bam.realign_hap) on fresh
    seeded cases -- no C-oracle hop in between (skipped only where oracle/_ref was not built),
  * a bounded slice of the live differential fuzz (tools/gpu_fuzz_live.py: random band radii / window sizes / time-slice
    lengths / max_n / max_l / gap penalties, tract-centred cases, multi-kb reads) on a seed that changes per round,
  * ALL 3,000 reads of BASELINE.json configs[1] against the C oracle (op strings, chunk scores, standardised CIGARs),
  * configs[4] at 8 Mb against a digest produced by the reference (tests/golden/make_golden_c5.py).
"""
import hashlib
import json
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

import oracle
from npore_b200 import cig, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROUND = 2          # seeds below derive from the round number: fresh cases every round


def _reference():
    import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref (the compiled reference) is not built on this box")
    return ref_loader.load_reference(6, 100, "/tmp/npore_ref_chain")


def test_cuda_vs_compiled_reference_direct(tables):
    """>= 300 fresh-seed cases incl. two-chunk 10 kb reads: CIGAR strings bit-exact, chunk scores bit-identical
    (north_star allows 1e-5 relative), straight against /root/reference/src/aln.pyx:379-787 as compiled."""
    from npore_b200.engine import Realigner
    ref = _reference()
    S, NP = tables
    ref.cfg.args.sub_scores, ref.cfg.args.np_scores = S, NP
    cm = synth.call_length_model(NP)
    rng = np.random.default_rng(9000 + ROUND)
    groups = {}
    for _ in range(320):                                        # small cases over the reference's own parameter grid
        rf, sq, cg, r, mb = synth.fuzz_case(rng, cm)
        groups.setdefault((r, mb), []).append((rf, sq, cg))
    big_ref, tr = synth.make_reference_with_tracts(200_000, rng)
    for rd in synth.make_reads(big_ref, 6, 10_400, rng, cm, tracts=tr):      # defaults: a 19,999-row chunk + a tail chunk
        groups.setdefault((30, 20000), []).append((rd[9], rd[7], cig.expand_cigar(rd[5])))
    n = two_chunk = 0
    for (r, mb), cases in groups.items():
        eng = Realigner(S, NP, r=r, max_b_rows=mb)
        refs = [oracle.bases_to_int(c[0]) for c in cases]; seqs = [oracle.bases_to_int(c[1]) for c in cases]
        outs, scores, status = eng.align_many(refs, seqs, [c[2] for c in cases])
        for k, c in enumerate(cases):
            want, wsc = ref.aln_sc.align(refs[k], seqs[k], c[2], S, NP, 5, 1, mb, r)
            assert outs[k] == want and status[k] == 0, f"r={r} max_b_rows={mb} case {k}: CIGAR differs from the compiled reference"
            assert np.array_equal(scores[k], np.asarray(wsc, np.float32)), f"r={r} max_b_rows={mb} case {k}: chunk scores differ"
            n += 1
            two_chunk += len(wsc) >= 2 and len(c[0]) > 9000
        eng.close()
    assert n >= 300 and two_chunk >= 4


def test_realign_read_records_vs_compiled_reference(tables, tmp_path):
    """bam.realign_read (bam.pyx:51-89) of the compiled reference against realign_reads here: whole SAM records."""
    from npore_b200 import bam as nbam, cfg
    ref = _reference()
    S, NP = tables
    ref.cfg.args.sub_scores, ref.cfg.args.np_scores = S, NP
    ref.cfg.args.out_prefix = str(tmp_path / "ref")
    cfg.args.sub_scores, cfg.args.np_scores = S, NP
    cfg.args.max_n, cfg.args.max_l, cfg.args.out_prefix = 6, 100, str(tmp_path / "gpu")
    rng = np.random.default_rng(9100 + ROUND)
    cm = synth.call_length_model(NP)
    rf, tr = synth.make_reference_with_tracts(60_000, rng)
    reads = synth.make_reads(rf, 24, 2500, rng, cm, tracts=tr)
    for rd in reads:
        ref.bam.realign_read(rd)
    want = open(str(tmp_path / "ref.sam")).read().splitlines()
    got = nbam.realign_reads(reads, write=False)
    assert got == want


def test_live_fuzz_slice(tables, monkeypatch, capsys):
    """tools/gpu_fuzz_live.py, 10 groups (~600 cases, every band-width template, odd radii, N bases, RR slices 7..100000)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gpu_fuzz_live
    monkeypatch.setenv("NPORE_RR_SLICE", "512"); monkeypatch.setenv("NPORE_STD_LONG_MIN", "4096")      # (the tool sets both per group; restored on teardown)
    monkeypatch.setattr(sys, "argv", ["gpu_fuzz_live.py", "10", str(77000 + ROUND), "1", "0"])
    rc = gpu_fuzz_live.main()
    out = capsys.readouterr().out
    assert "0 mismatches" in out and not rc, out[-2000:]


def _oracle_read(rd):
    S, NP = _TAB
    ir, iq = oracle.bases_to_int(rd[9]), oracle.bases_to_int(rd[7])
    o, sc, st = oracle.align(ir, iq, cig.expand_cigar(rd[5]), S, NP, return_scores=True)
    return (hashlib.sha256(o.encode()).hexdigest(), np.asarray(sc, np.float32).tobytes(), st,
            oracle.collapse_cigar(oracle.standardize(o, ir, iq)))


def _init_tab():
    global _TAB
    t = np.load(os.path.join(ROOT, "tests", "golden", "tables.npz"))
    _TAB = (t["sub_scores"], t["np_scores"])


def test_full_c2_every_read_vs_oracle(tables):
    """BASELINE.json configs[1] at full size: ALL 3,000 reads -- op strings, chunk scores, status and the standardised
    collapsed CIGAR (what the SAM record carries) -- against the C oracle run on the box's host cores."""
    import bench
    from npore_b200.engine import NPORE_OUT_RLE, NPORE_OUT_STANDARDIZE, Realigner
    S, NP = tables
    _, reads = bench.make_workload(20260101, 1_000_000, 3000, 10_000, NP)
    packed = bench.pack_reads(reads, pinned=False)
    eng = Realigner(S, NP)
    raw = eng.align_packed(packed, 0, eng.new_result(packed, 0, pinned=False))
    std = eng.align_packed(packed, NPORE_OUT_STANDARDIZE | NPORE_OUT_RLE, eng.new_result(packed, NPORE_OUT_STANDARDIZE | NPORE_OUT_RLE, pinned=False))
    std_txt = std.cigar_texts()
    cores = min(64, os.cpu_count() or 1)
    with mp.get_context("spawn").Pool(cores, initializer=_init_tab) as pool:
        want = pool.map(_oracle_read, reads, chunksize=max(1, len(reads) // (8 * cores)))
    bad = []
    for k, (h, sc, st, sd) in enumerate(want):
        ok = (hashlib.sha256(raw.ops[raw.ops_off[k]:raw.ops_off[k + 1]].tobytes()).hexdigest() == h and raw.scores(k).tobytes() == sc
              and int(raw.status[k]) == st and std_txt[k] == sd)
        if not ok:
            bad.append(k)
    assert not bad, f"{len(bad)} of {len(reads)} reads differ from the oracle, first: {bad[:5]}"
    eng.close()


def test_c5_8mb_against_reference_digest(tables):
    """BASELINE.json configs[4] at 8 Mb per haplotype (~800 chunks per item, `parts > 1` plan / finish kernels,
    standardize_long_kernel, global equality words): raw CIGAR, chunk scores and standardised CIGAR must hash to what the
    compiled reference produced (tests/golden/make_golden_c5.py)."""
    path = os.path.join(ROOT, "tests", "golden", "c5_8mb_digest.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/c5_8mb_digest.json not generated")
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_c5 import c5_haplotypes, sha
    from npore_b200.engine import NPORE_OUT_STANDARDIZE, PackedBatch, Realigner
    want = json.load(open(path))["haplotypes"]
    S, NP = tables
    haps = c5_haplotypes()
    packed = PackedBatch.from_strings([h[0] for h in haps], [h[1] for h in haps], [h[2] for h in haps])
    eng = Realigner(S, NP)
    raw = eng.align_packed(packed, 0, eng.new_result(packed, 0, pinned=False))
    std = eng.align_packed(packed, NPORE_OUT_STANDARDIZE, eng.new_result(packed, NPORE_OUT_STANDARDIZE, pinned=False))
    for h in (0, 1):
        assert sha(haps[h][2].encode()) == want[h]["input_cigar_sha"], "the workload generator no longer reproduces the digest's inputs"
        assert int(raw.status[h]) == 0 and len(raw.scores(h)) == want[h]["n_chunks"]
        assert sha(raw.ops[raw.ops_off[h]:raw.ops_off[h + 1]].tobytes()) == want[h]["raw"], f"haplotype {h}: raw CIGAR differs from the reference"
        assert sha(raw.scores(h).astype(np.float32).tobytes()) == want[h]["score"], f"haplotype {h}: chunk scores differ"
        assert sha(std.ops[std.ops_off[h]:std.ops_off[h + 1]].tobytes()) == want[h]["std"], f"haplotype {h}: standardised CIGAR differs"
    eng.close()
