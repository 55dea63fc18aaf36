"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): a few fuzz groups + multi-chunk reads."""
import gzip, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import oracle
from npore_b200 import synth
from npore_b200.engine import Realigner
t = np.load(os.path.join(ROOT, "tests/golden/tables.npz")); S, NP = t["sub_scores"], t["np_scores"]
fz = json.load(gzip.open(os.path.join(ROOT, "tests/golden/fuzz.json.gz"), "rt"))
bad = 0
for (r, mb) in [(30, 200), (10, 37), (30, 20000)]:
    cases = [c for c in fz if c["r"] == r and c["max_b_rows"] == mb][:8]
    os.environ["NPORE_RR_SLICE"] = "40"
    eng = Realigner(S, NP, max_b_rows=mb, r=r)
    refs = [oracle.bases_to_int(c["ref"]) for c in cases]; seqs = [oracle.bases_to_int(c["seq"]) for c in cases]
    outs, scores, status = eng.align_many(refs, seqs, [c["cigar"] for c in cases])
    std, _, _ = eng.align_many(refs, seqs, [c["cigar"] for c in cases], standardize=True, collapse=True)
    bad += sum(o != c["out"] or s != c["std"] for o, s, c in zip(outs, std, cases))
    eng.get_np_info(refs[0])
    eng.close()
print("sanitize_run mismatches:", bad)
sys.exit(1 if bad else 0)
