"""Golden vectors of the confusion-matrix / score-table path, from the UNMODIFIED compiled reference (oracle/_ref).

Run where /root/reference exists:   python tests/golden/make_golden_cm.py
  cm_guppy5.npz          the reference's shipped counts guppy5_stats/{subs,nps,inss,dels}_cm.npy; together with
                         tables.npz (calc_score_matrices of exactly these, aln.pyx:62-96) they pin aln.calc_score_matrices.
  confusion_kats.json.gz seeded alignments (npore_b200/synth.py:make_aligned_reads) + windows, and what the reference's
                         bam.calc_confusion_matrices (bam.pyx:351-510) returns for them when its `samtools mpileup` pipe
                         (bam.pyx:300-314; samtools is not installed here) is replaced by oracle/pileup_oracle.py's
                         restatement of the mpileup text.  nps is stored sparsely as [n-1, l, call, count]; qualities as phred+33 text.
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pileup_oracle as po, ref_loader  # noqa: E402
from npore_b200 import synth  # noqa: E402


def reference_counts(ref, contig, reads, start, end):
    lines = po.mpileup_column5([po.Read(*r) for r in reads], start, end)
    ref.bam.get_pileups = lambda bam, ctg, s, e: iter(lines)
    ref.bam.count_chunks = lambda regions: 1
    ref.cfg.args.refs = {"c": contig}
    ref.cfg.args.bam = None
    ref.cfg.args.regions = [("c", start, end)]
    ref.cfg.args.chunk_width = 100000
    return [np.asarray(m) for m in ref.bam.calc_confusion_matrices(("c", start, end))]


def main():
    ref = ref_loader.load_reference()
    d = "/root/reference/guppy5_stats/"
    np.savez_compressed(os.path.join(HERE, "cm_guppy5.npz"), **{k: np.load(f"{d}{k}_cm.npy") for k in ("subs", "nps", "inss", "dels")})
    cases = []
    for seed in range(24):
        rng = np.random.default_rng(7000 + seed)
        L = int(rng.integers(400, 2500))
        contig = synth.make_reference(L, rng, p_np=0.4)
        if seed % 4 == 1:
            contig = contig[:L // 2] + contig[L // 2:].lower()
        n_reads = int(rng.integers(2, 50)) if seed % 6 else 150
        reads = synth.make_aligned_reads(contig, n_reads, int(rng.integers(60, 700)), rng)
        start = int(rng.integers(0, L // 3))
        end = int(rng.integers(start + 1, L))          # end + 1 <= L: np_info[pos+1] stays inside the reference's array
        if seed % 5 == 0:
            start, end = 0, L - 1
        subs, nps, inss, dels = reference_counts(ref, contig, reads, start, end)
        nz = np.argwhere(nps)
        cases.append({"seed": 7000 + seed, "contig": contig, "start": start, "end": end,
                      "reads": [[r[0], [[int(n), op] for n, op in r[1]], r[2], None if r[3] is None else "".join(chr(q + 33) for q in r[3]), r[4], r[5]] for r in reads],
                      "subs": subs.tolist(), "inss": inss.tolist(), "dels": dels.tolist(),
                      "nps": [[int(a), int(b), int(c), int(nps[a, b, c])] for a, b, c in nz]})
    with gzip.open(os.path.join(HERE, "confusion_kats.json.gz"), "wt") as fh:
        json.dump(cases, fh)
    print(len(cases), "cases;", sum(len(c["reads"]) for c in cases), "reads")


if __name__ == "__main__":
    main()
