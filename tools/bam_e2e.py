"""BAM file in, realigned SAM file out: reads/s of npore_b200.bamio.realign_bam on the C2 workload (3,000 x 10 kb reads)
with the host time per phase (native BGZF/BAM decode, flat gather + submit, waiting for the GPU, native SAM formatting,
file write; the GPU runs concurrently with the host phases of the neighbouring batches),
next to the tuple API (get_read_data -> realign_reads), which is what a caller keeping the reference's per-read objects pays.
usage: python tools/bam_e2e.py [n_reads]"""
import os
import re
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from npore_b200 import bam as nbam, bamio, cfg  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    S, NP = bench.load_tables()
    cfg.args.sub_scores, cfg.args.np_scores = S, NP
    ref, reads = bench.make_workload(20260101, 1_000_000, n, 10000, NP)
    recs = [{"name": r[0], "flag": r[1], "ref_id": 0, "pos": r[3], "mapq": r[4], "cigar": [(int(a), b) for a, b in re.findall(r"(\d+)(\D)", r[5])],
             "seq": r[7], "qual": bytes([30] * len(r[7])), "tags": {"HP": r[10]}} for r in sorted(reads, key=lambda r: r[3])]
    t = time.perf_counter()
    bamio.write_bam("/tmp/c2.bam", "@HD\tVN:1.6\tSO:coordinate\n", [("chr1", len(ref))], recs)
    print(f"fixture BAM written in {time.perf_counter() - t:.1f} s ({os.path.getsize('/tmp/c2.bam') / 1e6:.1f} MB)", flush=True)
    fa = {"chr1": ref}
    bamio.realign_bam("/tmp/c2.bam", fa, out_prefix="/tmp/c2_out", argv=["bam_e2e"], max_reads=64)       # warm-up: context, kernels
    for rep in range(2):
        tm = {}
        t = time.perf_counter()
        got = bamio.realign_bam("/tmp/c2.bam", fa, out_prefix="/tmp/c2_out", argv=["bam_e2e"], timings=tm)
        dt = time.perf_counter() - t
        print(f"realign_bam: {got} reads in {dt:.3f} s = {got / dt:.0f} reads/s; phases (s): " + ", ".join(f"{k} {v:.3f}" for k, v in tm.items())
              + f"; SAM {os.path.getsize('/tmp/c2_out.sam') / 1e6:.1f} MB", flush=True)
    m = min(n, 300)
    t = time.perf_counter()
    tuples = list(bamio.get_read_data("/tmp/c2.bam", fa, max_reads=m))
    t1 = time.perf_counter()
    cfg.args.out_prefix = "/tmp/c2_out_tuples"
    nbam.realign_reads(tuples, write=True)
    t2 = time.perf_counter()
    print(f"tuple API on the first {m} reads: get_read_data {m / (t1 - t):.0f} reads/s, realign_reads {m / (t2 - t1):.0f} reads/s")
    a = [l for l in open("/tmp/c2_out.sam") if not l.startswith("@")][:m]
    b = open("/tmp/c2_out_tuples.sam").read().splitlines(True)[-m:]
    print("records identical between the two paths:", a == b)


if __name__ == "__main__":
    main()
