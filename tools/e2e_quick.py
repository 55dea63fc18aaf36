"""File -> file e2e (bamio.realign_bam defaults) on the C2 fixture: best and median of a few calls (tools/e2e_quick.py [n_reads] [reps])."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from npore_b200 import bamio, cfg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 9
S, NP = bench.load_tables()
cfg.args.sub_scores, cfg.args.np_scores = S, NP
ref, reads = bench.make_workload(20260101, 1_000_000, n, 10000, NP)
bench.write_fixture_bam("/tmp/c2.bam", ref, reads)
fa = {"chr1": ref}
ts = []
for rep in range(reps + 2):
    tm = {}
    t = time.perf_counter()
    bamio.realign_bam("/tmp/c2.bam", fa, out_prefix="/tmp/c2_out", argv=["x"], timings=tm)
    ts.append(time.perf_counter() - t)
ts = np.array(ts[2:])
print("all calls (ms):", " ".join(f"{1e3 * t:.1f}" for t in ts))
print(f"{n} reads: best {1e3 * ts.min():.1f} ms = {n / ts.min():.0f} reads/s, median {1e3 * np.median(ts):.1f} ms = {n / np.median(ts):.0f} reads/s  " +
      " ".join(f"{k} {1e3 * v:.0f}" for k, v in tm.items() if isinstance(v, float)))
