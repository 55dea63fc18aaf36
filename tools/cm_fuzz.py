"""GPU confusion matrices (npore_confusion_batch) vs the CPU oracle (oracle/pileup_oracle.py) on seeded random pileups.
usage: python tools/cm_fuzz.py [n_cases] [first_seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from npore_b200 import cfg, confusion, synth  # noqa: E402
from oracle import oracle, pileup_oracle as po  # noqa: E402


def case(seed):
    """{contig: (sequence, reads)}, [(contig, start, end)]: 1-2 contigs, 1-4 windows each, depth up to ~150."""
    rng = np.random.default_rng(seed)
    contigs, ranges = {}, []
    for name in ("c1", "c2")[:1 + seed % 2]:
        L = int(rng.integers(300, 4000))
        contig = synth.make_reference(L, rng, p_np=0.4)
        if seed % 4 == 1:
            contig = contig[:L // 2] + contig[L // 2:].lower()
        n_reads = int(rng.integers(1, 60)) if seed % 7 else int(rng.integers(200, 400))
        contigs[name] = (contig, synth.make_aligned_reads(contig, n_reads, int(rng.integers(50, 900)), rng))
        start = int(rng.integers(0, L // 3)); end = int(rng.integers(start + 1, L + 1))
        if seed % 5 == 0:
            start, end = 0, L
        cuts = sorted({start, end, *(int(x) for x in rng.integers(start, end + 1, size=int(rng.integers(0, 4))))})
        ranges += [(name, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    return contigs, ranges


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    s0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = 0
    for seed in range(s0, s0 + n):
        contigs, ranges = case(seed)
        want = None
        for c, a, b in ranges:
            one = po.confusion([po.Read(*r) for r in contigs[c][1]], contigs[c][0], a, b, oracle.get_np_info, oracle.bases_to_int)
            want = one if want is None else tuple(x + y for x, y in zip(want, one))
        got = confusion.calc_confusion_matrices_batch(ranges, refs={c: v[0] for c, v in contigs.items()},
                                                      reads={c: confusion.AlignedReads([r[:5] for r in v[1]]) for c, v in contigs.items()})
        if not all(np.array_equal(a, b) for a, b in zip(want, got)):
            bad += 1
            for nm, a, b in zip(("subs", "nps", "inss", "dels"), want, got):
                if not np.array_equal(a, b):
                    w = np.argwhere(a != b)
                    print(f"seed {seed} {nm}: {len(w)} cells differ, first {w[:4].tolist()} want {a[a != b][:4]} got {b[a != b][:4]}")
    print(f"{n} cases, {bad} mismatching")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
