"""Digest of BASELINE.json configs[4] ("C5", standardize_vcf path) at 8 Mb, produced by the UNMODIFIED compiled reference
(oracle/_ref; needs /root/reference -- run in the build container, commit the JSON):

    python tests/golden/make_golden_c5.py        ->  tests/golden/c5_8mb_digest.json

Workload (deterministic, npore_b200.synth; the same statements as tools/bench_configs.py): rng(20260105), an 8 Mb
n-polymer-rich contig, two haplotypes, each carrying copy-number changes (drawn from the learned call-length model) on a random
half of the contig's tracts plus 0.05 % substitutions; input CIGAR = the true edit script.  Per haplotype the digest holds
  raw   SHA-256 of the expanded CIGAR that aln.align returns (aln.pyx:379-787; ~800 chunks of 20,000 anti-diagonals)
  score SHA-256 of the float32 chunk scores (the 3-line score patch of oracle/build_ref.py)
  std   SHA-256 of the standardised expanded CIGAR that bam.realign_hap returns (bam.pyx:93-123)
The GPU test (tests/test_parity_chain.py::test_c5_8mb_against_reference_digest) must reproduce all three.
"""
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def c5_haplotypes(length=8_000_000, seed=20260105):
    from npore_b200 import synth
    t = np.load(os.path.join(ROOT, "tests", "golden", "tables.npz"))
    cm = synth.call_length_model(t["np_scores"])
    rng = np.random.default_rng(seed)
    ref, tr = synth.make_reference_with_tracts(length, rng)
    haps = []
    for _ in (1, 2):
        keep = tr[rng.random(len(tr)) < 0.5]
        seq, cg = synth.make_read(ref, rng, cm, p_ins=0.0, p_sub=0.0005, p_del=0.0, tracts=keep)
        haps.append((ref, seq, cg))
    return haps


def sha(b):
    return hashlib.sha256(b).hexdigest()


def work(job):
    kind, h = job
    import oracle
    import ref_loader
    ref = ref_loader.load_reference(6, 100, f"/tmp/c5_golden_{kind}_{h}")
    t = np.load(os.path.join(ROOT, "tests", "golden", "tables.npz"))
    S, NP = t["sub_scores"], t["np_scores"]
    ref.cfg.args.sub_scores, ref.cfg.args.np_scores = S, NP
    rf, sq, cg = c5_haplotypes()[h]
    t0 = time.time()
    if kind == "raw":
        out, sc = ref.aln_sc.align(oracle.bases_to_int(rf), oracle.bases_to_int(sq), cg, S, NP)
        return kind, h, {"raw": sha(out.encode()), "score": sha(np.asarray(sc, np.float32).tobytes()), "n_chunks": len(sc), "ops": len(out),
                         "seconds": round(time.time() - t0, 1)}
    res = ref.bam.realign_hap(("ctg", h + 1, sq, rf, cg))
    return kind, h, {"std": sha(res[4].encode()), "std_ops": len(res[4]), "seconds": round(time.time() - t0, 1)}


def main():
    jobs = [("raw", 0), ("raw", 1), ("std", 0), ("std", 1)]
    with mp.get_context("spawn").Pool(4) as pool:
        got = pool.map(work, jobs, chunksize=1)
    haps = c5_haplotypes()
    out = {"workload": "npore_b200.synth, rng(20260105), 8,000,000 bp contig, 2 haplotypes (see docstring)", "haplotypes": [{}, {}]}
    for kind, h, d in got:
        out["haplotypes"][h].update(d)
    for h in (0, 1):
        out["haplotypes"][h].update({"ref_len": len(haps[h][0]), "seq_len": len(haps[h][1]), "input_cigar_sha": sha(haps[h][2].encode())})
    with open(os.path.join(ROOT, "tests", "golden", "c5_8mb_digest.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
