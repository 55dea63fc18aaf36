"""File -> file e2e (bamio.realign_bam) on the C2 fixture for a few (window_bytes, max_batch_ops, n_inflight) settings."""
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
quick = os.environ.get("NPORE_SWEEP_QUICK")
for nin in ((int(x) for x in os.environ["NPORE_SWEEP_INFLIGHT"].split(",")) if os.environ.get("NPORE_SWEEP_INFLIGHT") else (2, 3)):
    for win in ((4 << 20, 6 << 20, 8 << 20, 12 << 20, 16 << 20) if quick else (0, 8 << 20, 16 << 20, 32 << 20)):
        for mbo in ((64_000_000,) if quick else (8_000_000, 16_000_000, 32_000_000, 64_000_000)):
            best, ph = 1e9, None
            for rep in range(4):
                tm = {}
                t = time.perf_counter()
                bamio.realign_bam("/tmp/c2.bam", fa, out_prefix="/tmp/c2_out", argv=["x"], timings=tm, window_bytes=win, max_batch_ops=mbo, n_inflight=nin)
                dt = time.perf_counter() - t
                if rep and dt < best:
                    best, ph = dt, tm
            print(f"inflight {nin} window {win >> 20:3d} MB batch_ops {mbo // 1_000_000:3d} M: {1e3 * best:6.1f} ms = {n / best:7.0f} reads/s  " +
                  " ".join(f"{k} {1e3 * v:.0f}" for k, v in ph.items() if isinstance(v, float)), flush=True)
