"""CPU differential fuzz of oracle/pull_model.c:pm2_align (the round-2 'INF ring' dataflow of csrc/forward.cuh) against the
C oracle (oracle/npore_oracle.c).  usage: pm2_fuzz.py [n_cases] [seed]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402
from npore_b200 import synth  # noqa: E402


def load():
    here = os.path.join(ROOT, "oracle")
    subprocess.check_call(["make", "-C", here, "libpull_model.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(here, "libpull_model.so"))
    u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L.pm2_align.restype = C.c_int64
    L.pm2_align.argtypes = [u8, C.c_int, u8, C.c_int, C.c_char_p, C.c_int64, f32, f32, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                            C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int64, f32, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                            C.POINTER(C.c_int64)]
    return L


def pm2(L, ir, iq, cg, S, NP, mb, r, form=0, max_n=6, max_l=100):
    pad = lambda a: a if a.size else np.zeros(1, np.uint8)  # noqa: E731
    cap = len(ir) + len(iq) + 8
    out = C.create_string_buffer(cap)
    sc = np.zeros(cap // max(1, mb - 1) + 4, np.float32)
    ns, st, nr = C.c_int(0), C.c_int(0), C.c_int64(0)
    cb = cg.encode()
    n = L.pm2_align(pad(ir), len(ir), pad(iq), len(iq), cb, len(cb), S, NP, NP.shape[1], max_n, max_l, 5.0, 1.0, mb, r, form, out, cap,
                    sc, len(sc), C.byref(ns), C.byref(st), C.byref(nr))
    return out.raw[:n].decode(), sc[:ns.value].copy(), st.value, nr.value


FORMS = (0, 1)


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    t = np.load(os.path.join(ROOT, "tests", "golden", "tables.npz"))
    S, NP = t["sub_scores"], t["np_scores"]
    cm = synth.call_length_model(NP)
    rng = np.random.default_rng(seed)
    L = load()
    bad = risky = 0
    for k in range(n_cases):
        rf, sq, cg, r, mb = synth.fuzz_case(rng, cm)
        if rng.random() < 0.5:
            r = int(rng.choice([13, 14, 15, 29, 30, 31, 62, 63]))
        if rng.random() < 0.3:      # long INDEL runs in the input path (aliasing corner)
            pos = int(rng.integers(0, len(cg) + 1))
            if rng.random() < 0.5:
                ins = "".join(rng.choice(list("ACGT"), size=int(rng.integers(6, 12))))
                # insert bases into the read at the matching read offset
                ro = sum(1 for c in cg[:pos] if c != "D")
                sq = sq[:ro] + ins + sq[ro:]
                cg = cg[:pos] + "I" * len(ins) + cg[pos:]
            else:
                fo = sum(1 for c in cg[:pos] if c != "I")
                dele = "".join(rng.choice(list("ACGT"), size=int(rng.integers(6, 12))))
                rf = rf[:fo] + dele + rf[fo:]
                cg = cg[:pos] + "D" * len(dele) + cg[pos:]
        ir, iq = oracle.bases_to_int(rf), oracle.bases_to_int(sq)
        want, wsc, wst = oracle.align(ir, iq, cg, S, NP, max_b_rows=mb, r=r, return_scores=True)
        for form in FORMS:
            got, gsc, gst, nr = pm2(L, ir, iq, cg, S, NP, mb, r, form)
            risky += nr
            if got != want or gst != wst or not np.array_equal(gsc, wsc):
                bad += 1
                print(f"MISMATCH case {k} form {form} r={r} mb={mb} len={len(ir)}/{len(iq)}")
                break
    print(f"{n_cases} cases, {bad} mismatches, {risky} risky anti-diagonals took the checked path")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
