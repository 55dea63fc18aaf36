"""Development check (run under gpurun): CUDA path vs the C oracle on seeded fuzz + golden fixtures.
Writes a report to gpurun_out/fuzz_report.txt.  (The parity tests proper live in tests/.)"""
import gzip, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import oracle
from npore_b200 import synth
from npore_b200.engine import Realigner

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
rep = open(os.path.join(ROOT, "gpurun_out", "fuzz_report.txt"), "w")
def log(*a):
    s = " ".join(str(x) for x in a); print(s); rep.write(s + "\n"); rep.flush()

t = np.load(os.path.join(ROOT, "tests/golden/tables.npz")); S, NP = t["sub_scores"], t["np_scores"]
fz = json.load(gzip.open(os.path.join(ROOT, "tests/golden/fuzz.json.gz"), "rt"))
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else len(fz)
groups = {}
for c in fz[:n_cases]:
    groups.setdefault((c["r"], c["max_b_rows"]), []).append(c)
bad = 0; total = 0
for (r, mb), cases in sorted(groups.items()):
    t0 = time.time()
    eng = Realigner(S, NP, max_b_rows=mb, r=r)
    refs = [oracle.bases_to_int(c["ref"]) for c in cases]; seqs = [oracle.bases_to_int(c["seq"]) for c in cases]
    outs, scores, status = eng.align_many(refs, seqs, [c["cigar"] for c in cases])
    std, _, _ = eng.align_many(refs, seqs, [c["cigar"] for c in cases], standardize=True, collapse=True)
    nb = 0
    for k, c in enumerate(cases):
        total += 1
        ok = outs[k] == c["out"] and np.array_equal(scores[k], np.array(c["scores"], np.float32)) and std[k] == c["std"] and status[k] == 0
        if not ok:
            bad += 1; nb += 1
            if nb <= 3:
                log(f"MISMATCH r={r} mb={mb} case={k} Lr={len(c['ref'])} Ls={len(c['seq'])} status={status[k]}")
                log("  ref ", c["ref"][:120]); log("  seq ", c["seq"][:120]); log("  cig ", c["cigar"][:120])
                log("  want", c["out"][:120], c["scores"][:4]); log("  got ", outs[k][:120], scores[k][:4])
                log("  wstd", c["std"][:100]); log("  gstd", std[k][:100])
    # np_info
    for c in cases[:20]:
        if len(c["ref"]):
            a = eng.get_np_info(oracle.bases_to_int(c["ref"])); b = oracle.get_np_info(oracle.bases_to_int(c["ref"]))
            if not np.array_equal(a, b):
                bad += 1; log("NP_INFO mismatch", c["ref"][:80])
    log(f"group r={r} mb={mb}: {len(cases)} cases, {nb} bad, {time.time()-t0:.2f}s, stats={ {k: v for k, v in eng.stats().items() if k.startswith('ms_') or k in ('n_chunks','launches')} }")
    eng.close()
log(f"TOTAL {total} cases, {bad} bad")
sys.exit(1 if bad else 0)
