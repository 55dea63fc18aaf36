"""Timeline of one bamio.realign_bam call on the C2 fixture (seconds since the call, event)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from npore_b200 import bamio, cfg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
S, NP = bench.load_tables()
cfg.args.sub_scores, cfg.args.np_scores = S, NP
ref, reads = bench.make_workload(20260101, 1_000_000, n, 10000, NP)
bench.write_fixture_bam("/tmp/c2.bam", ref, reads)
fa = {"chr1": ref}
for rep in range(int(os.environ.get('NPORE_REPS', '3'))):
    tm = {"trace": []}
    t = time.perf_counter()
    bamio.realign_bam("/tmp/c2.bam", fa, out_prefix="/tmp/c2_out", argv=["x"], timings=tm)
    dt = time.perf_counter() - t
print(f"total {1e3 * dt:.1f} ms")
for t, ev in tm["trace"]:
    print(f"{1e3 * t:8.2f} ms  {ev}")
