"""Where does bam.realign_reads spend its time? (host packing / GPU / SAM formatting) -- run under gpurun."""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from npore_b200 import bam, cfg
S, NP = bench.load_tables()
cfg.args.sub_scores, cfg.args.np_scores = S, NP
cfg.args.out_prefix = "/tmp/npore_glue"
_, reads = bench.make_workload(20260101, 1_000_000, 3000, 10000, NP)
bam.realign_reads(reads[:64], write=False)        # warm up (context, kernels)
if os.path.exists("/tmp/npore_glue.sam"): os.remove("/tmp/npore_glue.sam")
t0 = time.perf_counter(); lines = bam.realign_reads(reads, write=True); dt = time.perf_counter() - t0
print(f"realign_reads: {len(reads)} reads in {dt:.3f} s = {len(reads)/dt:.0f} reads/s; SAM bytes {sum(len(l) for l in lines)}")
pr = cProfile.Profile(); pr.enable(); bam.realign_reads(reads, write=True); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
