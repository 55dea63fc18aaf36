"""BAM file in -> SAM file out (bamio.realign_bam: native ingest, GPU, native SAM text) on random BAMs against the oracle:
reads with soft / hard clips, missing qualities, a soft-masked (lower-case) reference, HP tags, secondary / supplementary / unmapped
records to skip, two contigs, streaming windows and batch sizes chosen at random.  usage: python tools/bam_fuzz.py [n_files] [seed]"""
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402
from npore_b200 import bamio, cfg, synth  # noqa: E402


def main():
    n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    t = np.load(os.path.join(ROOT, "tests/golden/tables.npz")); S, NP = t["sub_scores"], t["np_scores"]
    cfg.args.sub_scores, cfg.args.np_scores = S, NP
    cm = synth.call_length_model(NP)
    rng = np.random.default_rng(seed)
    bad = total = 0
    with tempfile.TemporaryDirectory() as d:
        for f in range(n_files):
            contigs = {}
            recs = []
            expect = []
            for ci, name in enumerate(("chrA", "chrB")[:1 + f % 2]):
                ref, tr = synth.make_reference_with_tracts(int(rng.integers(5000, 40000)), rng)
                if f % 3 == 0:
                    ref = ref[:len(ref) // 2] + ref[len(ref) // 2:].lower()
                contigs[name] = ref
                for rd in synth.make_reads(ref.upper(), int(rng.integers(1, 40)), int(rng.integers(200, 3000)), rng, cm, tracts=tr):
                    name_, flag, _, start, mapq, cigar, stop, seq, quals, rref, hap = rd
                    groups = [(int(a), b) for a, b in re.findall(r"(\d+)(\D)", cigar)]
                    lead = "".join(rng.choice(list("ACGT"), size=int(rng.integers(0, 6)))) if rng.random() < 0.4 else ""
                    trail = "".join(rng.choice(list("ACGT"), size=int(rng.integers(0, 6)))) if rng.random() < 0.4 else ""
                    cg = ([(int(rng.integers(1, 9)), "H")] if rng.random() < 0.2 else []) + ([(len(lead), "S")] if lead else []) + groups + \
                         ([(len(trail), "S")] if trail else [])
                    body = seq
                    flag = int(rng.choice([0, 16, 0, 16, 256, 2048, 4, 1024]))
                    has_q = rng.random() < 0.8
                    full = lead + body + trail
                    rec = {"name": f"{name}_{len(recs)}", "flag": flag, "ref_id": ci if flag != 4 else -1, "pos": start if flag != 4 else -1, "mapq": int(rng.integers(0, 61)),
                           "cigar": cg if flag != 4 else [], "seq": full, "qual": bytes(rng.integers(0, 50, size=len(full)).astype(np.uint8)) if has_q else None,
                           "tags": {"HP": int(rng.integers(1, 3))} if rng.random() < 0.5 else {}}
                    recs.append(rec)
                    if not flag & (0x4 | 0x100 | 0x800):
                        want = oracle.realign_cigar(ref.upper()[start:stop], seq.upper(), cigar, S, NP)
                        q = "*" if not has_q or not len(seq) else "".join(chr(33 + x) for x in rec["qual"][len(lead):len(lead) + len(seq)])
                        expect.append((ci, start, f"{rec['name']}\t{flag}\t{name}\t{start + 1}\t{rec['mapq']}\t{want}\t*\t0\t{stop - start}\t{seq.upper()}\t{q}\tHP:i:{rec['tags'].get('HP', 0)}"))
            order = sorted(range(len(recs)), key=lambda k: (recs[k]["ref_id"] if recs[k]["ref_id"] >= 0 else 99, recs[k]["pos"]))
            recs = [recs[k] for k in order]
            expect = [e[2] for e in sorted(expect, key=lambda e: (e[0], e[1]))]
            bam = os.path.join(d, f"f{f}.bam")
            bamio.write_bam(bam, "@HD\tVN:1.6\tSO:coordinate\n", [(n, len(s)) for n, s in contigs.items()], recs)
            n = bamio.realign_bam(bam, contigs, out_prefix=os.path.join(d, f"o{f}"), argv=["bam_fuzz"], window_bytes=int(rng.choice([2000, 50000, 64 << 20])),
                                  max_batch_ops=int(rng.choice([5000, 200000, 64_000_000])), n_inflight=int(rng.integers(1, 4)))
            got = [l for l in open(os.path.join(d, f"o{f}.sam")).read().splitlines() if not l.startswith("@")]
            # records with equal (contig, pos) may come in either order of the stable sort above: compare as sorted lists per key
            total += len(expect)
            if n != len(expect) or sorted(got) != sorted(expect) or [g.split("\t")[3] for g in got] != [e.split("\t")[3] for e in expect]:
                bad += 1
                print(f"MISMATCH file {f}: {n} written, {len(expect)} expected, first differing record: "
                      f"{next((a[:80] + ' <> ' + b[:80] for a, b in zip(sorted(got), sorted(expect)) if a != b), 'count')}", flush=True)
    print(f"bam fuzz seed {seed}: {n_files} files, {total} records, {bad} files mismatching")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
