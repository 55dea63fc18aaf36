"""Throughput of the confusion-matrix path (npore_confusion_batch) on a C2-shaped pileup: 1 Mb contig, 3,000 x 10 kb reads,
chunk_width 100,000 (10 windows).  Prints device-call time (H2D + kernels + D2H through the C ABI), host packing time, and --
on a sample window -- the CPU time of the reference's own parser (compiled bam.calc_confusion_matrices when oracle/_ref is
present, else the oracle port) on ready-made pileup text, i.e. without the samtools process the reference also pays for.
usage: python tools/cm_bench.py [n_reads] [read_len] [contig_len]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from npore_b200 import cfg, confusion, synth  # noqa: E402
import oracle  # noqa: E402
import pileup_oracle as po  # noqa: E402
import ref_loader  # noqa: E402


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    read_len = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
    rng = np.random.default_rng(20260101)
    contig = synth.make_reference(L, rng, p_np=0.3)
    t = time.time()
    reads = synth.make_aligned_reads(contig, n_reads, read_len, rng, with_clips=False)
    print(f"generated {n_reads} reads in {time.time() - t:.1f}s", flush=True)
    t = time.time()
    ar = confusion.AlignedReads([r[:5] for r in reads])
    t_flat = time.time() - t
    cfg.args.chunk_width = 100000
    ranges = confusion.get_ranges([("c", 0, L)])
    eng = confusion._np_engine()
    t = time.time()
    pack = confusion.PileupPack(ranges, {"c": contig}, {"c": ar}, 6)
    t_pack = time.time() - t
    eng.confusion_batch(pack)
    ts = []
    for _ in range(5):
        t = time.time(); out = eng.confusion_batch(pack); ts.append(time.time() - t)
    entries = int((ar.end - ar.pos).sum())
    dt = float(np.median(ts))
    print(f"GPU: {dt * 1e3:.1f} ms per call ({entries / dt / 1e6:.0f} M pileup entries/s, {L / dt / 1e6:.1f} Mb/s of reference at depth "
          f"{entries / L:.0f}); flatten {t_flat:.2f}s, pack {t_pack * 1e3:.0f} ms; base entries counted {int(out[0].sum())}")
    # CPU parser on one 20 kb window
    s, e = L // 2, L // 2 + 20000
    rl = [po.Read(*r) for r in reads if r[0] < e]
    lines = po.mpileup_column5(rl, s, e)
    chars = sum(len(x) for x in lines)
    if ref_loader.available():
        ref = ref_loader.load_reference()
        ref.bam.get_pileups = lambda bam, ctg, a, b: iter(lines)
        ref.bam.count_chunks = lambda regions: 1
        ref.cfg.args.refs = {"c": contig}; ref.cfg.args.bam = None; ref.cfg.args.regions = [("c", s, e)]; ref.cfg.args.chunk_width = 100000
        t = time.time(); want = [np.asarray(m) for m in ref.bam.calc_confusion_matrices(("c", s, e))]; tc = time.time() - t
        kind = "reference"
    else:
        info = oracle.get_np_info(oracle.bases_to_int(contig[s:e + 1]))
        t = time.time(); want = po.confusion_from_lines(lines, contig, s, e, info); tc = time.time() - t
        kind = "port"
    got = eng.confusion_batch(confusion.PileupPack([("c", s, e)], {"c": contig}, {"c": ar}, 6))
    same = all(np.array_equal(a, b) for a, b in zip(want, got))
    n_e = int(want[0].sum())
    print(f"\nCPU ({kind}, 1 core, text ready): {tc:.2f}s for a 20 kb window ({chars} chars, {n_e / tc / 1e6:.2f} M base entries/s); "
          f"GPU result identical: {same}")


if __name__ == "__main__":
    main()
