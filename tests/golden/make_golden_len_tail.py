"""Golden vectors for LEN candidates on the LAST copies of a read tract (an inserted run of tract units followed by a
different base), from the UNMODIFIED compiled reference (oracle/_ref; aln_sc = the build-time copy that also returns the
chunk scores).  Found by the fresh-seed differential fuzz (tools/gpu_fuzz_live.py): aln.pyx:606-607 compares
seq[i-n .. i) with ref[j .. j+n), and on the last copies of a read tract seq[i .. i+n) is no longer the same unit.

Run where /root/reference exists:   python tests/golden/make_golden_len_tail.py   ->  tests/golden/len_tail_kats.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_loader  # noqa: E402
from npore_b200 import synth  # noqa: E402


def main():
    ref = ref_loader.load_reference()
    S, NP = ref_loader.reference_tables(ref)
    cases = [("CAAAAAAAAAAAAAAAA", "CAAAATAAAAAAAAAAAAAAAA", 8, 1), ("CAAAAAAAAAAAAA", "CAAAAATAAAAAAAAAAAAA", 8, 1),
             ("GC" + "A" * 45, "GCAAAC" + "A" * 45, 8, 1), ("CGTAGAATACCTC" + "A" * 18, "TCAAAACAAAAAA", 8, 1),
             ("GATC" + "AC" * 9 + "GT", "GATC" + "ACACACT" + "AC" * 9 + "GT", 10, 6), ("TTG" + "CAG" * 7 + "TT", "TTG" + "CAGCAGCAGA" + "CAG" * 7 + "TT", 30, 6)]
    rng = np.random.default_rng(4242)
    for _ in range(150):
        hp = int(rng.integers(8, 130))
        unit = "".join(rng.choice(list("ACGT"), size=int(rng.choice([1, 1, 1, 2, 3]))))
        if len(set(unit)) == 1:
            unit = unit[0]
        pre = "".join(rng.choice(list("ACGT"), size=int(rng.integers(4, 30)))); suf = "".join(rng.choice(list("ACGT"), size=int(rng.integers(4, 30))))
        copies = max(3, hp // len(unit))
        rf = pre + unit * copies + suf
        k = int(rng.integers(max(3, copies - 12), copies + 12))
        sq0 = pre + unit * k + suf
        sq, _ = synth.make_read(sq0, rng, None, p_ins=0.03, p_sub=0.03, p_del=0.03)
        cases.append((rf, sq, int(rng.choice([5, 8, 16, 30])), int(rng.choice([6, 6, 3, 1]))))
    out = []
    for rf, sq, r, max_n in cases:
        ref.cfg.args.max_n = max_n
        m = min(len(rf), len(sq))
        cigar = "M" * m + "D" * (len(rf) - m) + "I" * (len(sq) - m)
        ir, iq = ref.cig.bases_to_int(rf), ref.cig.bases_to_int(sq)
        aln, scores = ref.aln_sc.align(ir, iq, cigar, S, NP, 5, 1, 20000, r)
        out.append({"ref": rf, "seq": sq, "cigar": cigar, "r": r, "max_n": max_n, "out": aln, "scores": [float(np.float32(x)) for x in scores]})
    ref.cfg.args.max_n = 6
    json.dump(out, open(os.path.join(HERE, "len_tail_kats.json"), "w"))
    print(len(out), "cases")


if __name__ == "__main__":
    main()
