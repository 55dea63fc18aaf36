"""Forward-kernel time of C2-like batches of n reads under each kernel form (NPORE_TEAM), for the team heuristic of api.cu
(tools/batch_size_sweep.py [n ...])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from npore_b200.engine import NPORE_OUT_NO_EXPANDED, NPORE_OUT_RLE, NPORE_OUT_STANDARDIZE, Realigner
S, NP = bench.load_tables()
flags = NPORE_OUT_STANDARDIZE | NPORE_OUT_RLE | NPORE_OUT_NO_EXPANDED
sizes = [int(x) for x in sys.argv[1:]] or [100, 250, 500, 750, 1000, 1500, 2000, 3000]
_, reads = bench.make_workload(20260101, 1_000_000, max(sizes), 10000, NP)
print("| reads | team | forward ms | kernels ms | GCUPS (kernels) |\n|---|---|---|---|---|")
for n in sizes:
    packed = bench.pack_reads(reads[:n], pinned=False)
    for team in ("1", "2", ""):
        if team:
            os.environ["NPORE_TEAM"] = team
        else:
            os.environ.pop("NPORE_TEAM", None)
        eng = Realigner(S, NP)
        eng.upload(packed)
        best = None
        for _ in range(4):
            eng.run(flags)
            st = eng.stats()
            if best is None or st["ms_forward"] < best["ms_forward"]:
                best = st
        print(f"| {n} | {team or 'auto'} | {best['ms_forward']:.2f} | {best['ms_kernels_total']:.2f} | {best['n_cu'] / best['ms_kernels_total'] / 1e6:.1f} |", flush=True)
        eng.close()
