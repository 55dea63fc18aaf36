"""File -> file on several GPUs (bamio.realign_bam(devices=[...]), region sharding): a coordinate-sorted BAM of T tiles of the C2
generator as T contigs (3,000 reads each), realigned with 1 / 2 / 4 / 8 devices; the merged SAM of every run must be
byte-identical to the single-GPU file.  usage: python tools/e2e_multi.py [tiles=8] [reps=3]"""
import hashlib, os, re, sys, time
from multiprocessing import get_context
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from npore_b200 import bamio, cfg


def tile(t):
    S, NP = bench.load_tables()
    ref, reads = bench.make_workload(20260101 + 7919 * (t + 1), 1_000_000, 3000, 10000, NP)
    recs = [{"name": f"t{t}_{r[0]}", "flag": r[1], "ref_id": t, "pos": r[3], "mapq": r[4], "seq": r[7], "qual": bytes([30] * len(r[7])),
             "cigar": [(int(a), b) for a, b in re.findall(r"(\d+)(\D)", r[5])], "tags": {"HP": r[10]}} for r in sorted(reads, key=lambda r: r[3])]
    return ref, recs


if __name__ == "__main__":
    import torch
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    S, NP = bench.load_tables()
    cfg.args.sub_scores, cfg.args.np_scores = S, NP
    t0 = time.perf_counter()
    with get_context("spawn").Pool(min(T, os.cpu_count() or 1)) as pool:
        tiles = pool.map(tile, range(T))
    fa = {f"chr{t + 1}": tiles[t][0] for t in range(T)}
    recs = [r for t in range(T) for r in tiles[t][1]]
    bamio.write_bam("/tmp/multi.bam", "@HD\tVN:1.6\tSO:coordinate\n", [(f"chr{t + 1}", 1_000_000) for t in range(T)], recs, level=1)
    n = len(recs)
    print(f"{n} reads on {T} contigs, BAM {os.path.getsize('/tmp/multi.bam') / 1e6:.1f} MB, built in {time.perf_counter() - t0:.0f} s, "
          f"{os.cpu_count()} host cores, {torch.cuda.device_count()} GPUs", flush=True)
    want = None
    for G in (1, 2, 4, 8):
        if G > torch.cuda.device_count():
            break
        best, tm_best = 1e9, None
        for rep in range(reps + 1):
            tm = {}
            t1 = time.perf_counter()
            got = bamio.realign_bam("/tmp/multi.bam", fa, out_prefix=f"/tmp/multi_out{G}", argv=["x"], timings=tm, devices=list(range(G)) if G > 1 else None)
            dt = time.perf_counter() - t1
            if rep and dt < best:
                best, tm_best = dt, tm
        sha = hashlib.sha256(open(f"/tmp/multi_out{G}.sam", "rb").read()).hexdigest()
        want = want or sha
        print(f"| {G} | {1e3 * best:.0f} ms | {got / best:,.0f} reads/s | " + " ".join(f"{k} {1e3 * v:.0f}" for k, v in tm_best.items() if isinstance(v, float)) +
              f" | {'identical to 1 GPU' if sha == want else 'DIFFERS from 1 GPU'} |", flush=True)
