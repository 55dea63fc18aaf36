import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
from npore_b200 import synth
from npore_b200.engine import NPORE_OUT_RLE, NPORE_OUT_STANDARDIZE, NPORE_OUT_NO_EXPANDED, PackedBatch, Realigner
t = np.load("tests/golden/tables.npz"); S, NP = t["sub_scores"], t["np_scores"]
cm = synth.call_length_model(NP)
Lc = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
rng = np.random.default_rng(20260105)
t0 = time.time()
ref, tr = synth.make_reference_with_tracts(Lc, rng)
haps = []
for hap in (1, 2):
    keep = tr[rng.random(len(tr)) < 0.5]
    seq, cg = synth.make_read(ref, rng, cm, p_ins=0.0, p_sub=0.0005, p_del=0.0, tracts=keep)
    haps.append((ref, seq, cg))
print("generated in", round(time.time() - t0, 1), "s", flush=True)
eng = Realigner(S, NP)
packed = PackedBatch.from_strings([c[0] for c in haps], [c[1] for c in haps], [c[2] for c in haps])
for flags in (NPORE_OUT_STANDARDIZE | NPORE_OUT_RLE | NPORE_OUT_NO_EXPANDED, NPORE_OUT_STANDARDIZE):
    res = eng.new_result(packed, flags)
    for _ in range(2):
        t0 = time.perf_counter(); eng.align_packed(packed, flags, res); dt = time.perf_counter() - t0
    st = eng.stats()
    print(flags, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items() if k.startswith("ms_") or k in ("n_chunks", "n_sub_batches", "launches")}, "wall ms", round(dt * 1e3, 1), "GCUPS", round(st["n_cu"] / dt / 1e9, 1), "groups", int(res.rle_off[2]) if flags & 2 else "-", flush=True)
