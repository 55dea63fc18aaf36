"""One upload + a few runs of the hot path on a C2-like batch, for ncu (tools/prof_run.py [n_reads] [runs])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from npore_b200.engine import NPORE_OUT_NO_EXPANDED, NPORE_OUT_RLE, NPORE_OUT_STANDARDIZE, Realigner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S, NP = bench.load_tables()
_, reads = bench.make_workload(20260101, 1_000_000, n, 10000, NP)
packed = bench.pack_reads(reads, pinned=False)
eng = Realigner(S, NP)
flags = NPORE_OUT_STANDARDIZE | NPORE_OUT_RLE | NPORE_OUT_NO_EXPANDED
eng.upload(packed)
for _ in range(runs):
    eng.run(flags)
    print({k: round(v, 3) if isinstance(v, float) else v for k, v in eng.stats().items()})
