"""The BASELINE.json configurations other than the bench.py headline (C2): C1 golden reads (latency), C4 long reads with the
band radius x max_b_rows sweep, C5 whole-contig haplotypes.  One GPU; prints a markdown table:
cell updates, chunks, kernels-only and end-to-end (host buffers in, host buffers out) time, GCUPS, traceback bytes in
flight (b_rows * 32*TBS * 2 B per chunk, summed over the resident sub-batch) and resident forward warps per SM.
Parity at these configurations is covered by tests/test_gpu_parity.py; a sample item per row is re-checked against the
oracle here when --check is given.
usage: python tools/bench_configs.py [--quick] [--check]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from npore_b200 import cig, synth  # noqa: E402
from npore_b200.engine import NPORE_OUT_RLE, NPORE_OUT_STANDARDIZE, NPORE_OUT_NO_EXPANDED, PackedBatch, Realigner  # noqa: E402


def run(eng, cases, flags, reps=3):
    packed = PackedBatch.from_strings([c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases])
    res = eng.new_result(packed, flags)
    eng.align_packed(packed, flags, res)
    best_k, best_e = 1e30, 1e30
    for _ in range(reps):
        t = time.perf_counter()
        eng.align_packed(packed, flags, res)
        best_e = min(best_e, time.perf_counter() - t)
        st = eng.stats()
        best_k = min(best_k, st["ms_kernels_total"] / 1e3)
    return st, best_k, best_e, res


def row(name, st, tk, te):
    return (f"| {name} | {st['n_items']} | {st['n_chunks']} | {st['n_cu'] / 1e9:.2f} | {tk * 1e3:.1f} | {st['n_cu'] / tk / 1e9:.1f} | "
            f"{te * 1e3:.1f} | {st['n_cu'] / te / 1e9:.1f} | {st['tb_bytes'] / 2**30:.2f} | {st['n_sub_batches']} | {st['fwd_warps_per_sm']} |")


def main():
    quick = "--quick" in sys.argv
    check = "--check" in sys.argv
    t = np.load(os.path.join(ROOT, "tests", "golden", "tables.npz"))
    S, NP = t["sub_scores"], t["np_scores"]
    cm = synth.call_length_model(NP)
    if check:
        import oracle
    print("| config | items | chunks | G cell updates | kernels ms | GCUPS (kernels) | e2e ms | GCUPS (e2e) | traceback GiB | sub-batches | fwd warps/SM |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    # ---- C1: the reference's golden fixture, 10 short reads -> latency
    with open(os.path.join(ROOT, "tests", "golden", "golden_sam.json")) as fh:
        g = json.load(fh)
    cases = [(r[9], r[7], cig.expand_cigar(r[5])) for r in g["reads"]]
    eng = Realigner(S, NP)
    st, tk, te, _ = run(eng, cases, NPORE_OUT_STANDARDIZE | NPORE_OUT_RLE | NPORE_OUT_NO_EXPANDED, reps=10)
    print(row("C1 golden 10 reads (r=30)", st, tk, te), flush=True)
    eng.close()
    # ---- C4: 50-100 kb reads
    rng = np.random.default_rng(20260104)
    n4 = 64 if quick else 512
    ref, tr = synth.make_reference_with_tracts(4_000_000, rng)
    lens = rng.integers(50_000, 100_001, size=n4)
    reads = []
    for L in lens:
        reads += synth.make_reads(ref, 1, int(L), rng, cm, tracts=tr)
    cases = [(rd[9], rd[7], cig.expand_cigar(rd[5])) for rd in reads]
    radii = tuple(int(x) for x in os.environ["NPORE_CFG_R"].split(",")) if os.environ.get("NPORE_CFG_R") else (10, 30, 60, 100)
    for r in radii:
        for mb in (5000, 20000, 50000):
            eng = Realigner(S, NP, r=r, max_b_rows=mb)
            st, tk, te, res = run(eng, cases, NPORE_OUT_STANDARDIZE | NPORE_OUT_RLE | NPORE_OUT_NO_EXPANDED, reps=2)
            print(row(f"C4 {n4} reads 50-100 kb, r={r}, max_b_rows={mb}", st, tk, te), flush=True)
            if check:
                ir, iq = oracle.bases_to_int(cases[0][0]), oracle.bases_to_int(cases[0][1])
                want = oracle.collapse_cigar(oracle.standardize(oracle.align(ir, iq, cases[0][2], S, NP, max_b_rows=mb, r=r), ir, iq))
                assert res.cigar_text(0) == want, (r, mb)
            eng.close()
    # ---- C5: two haplotypes of one contig, indels inside tracts
    Lc = 2_000_000 if quick else 8_000_000
    rng = np.random.default_rng(20260105)
    ref, tr = synth.make_reference_with_tracts(Lc, rng)
    haps = []
    for hap in (1, 2):
        keep = tr[rng.random(len(tr)) < 0.5]
        seq, cg = synth.make_read(ref, rng, cm, p_ins=0.0, p_sub=0.0005, p_del=0.0, tracts=keep)
        haps.append((ref, seq, cg))
    eng = Realigner(S, NP)
    st, tk, te, res = run(eng, haps, NPORE_OUT_STANDARDIZE | NPORE_OUT_RLE | NPORE_OUT_NO_EXPANDED, reps=2)
    print(row(f"C5 2 haplotypes x {Lc / 1e6:.0f} Mb (r=30)", st, tk, te), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
