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
# (r, max_b_rows, NPORE_TEAM): few chunks pick the two-warp team kernels by themselves; "1" forces one chunk per warp
for (r, mb, team) in [(30, 200, "1"), (30, 200, None), (10, 37, None), (30, 20000, "1"), (30, 20000, None)]:
    cases = [c for c in fz if c["r"] == r and c["max_b_rows"] == mb][:8]
    os.environ["NPORE_RR_SLICE"] = "40"
    os.environ.pop("NPORE_TEAM", None)
    if team:
        os.environ["NPORE_TEAM"] = team
    eng = Realigner(S, NP, max_b_rows=mb, r=r)
    refs = [oracle.bases_to_int(c["ref"]) for c in cases]; seqs = [oracle.bases_to_int(c["seq"]) for c in cases]
    outs, scores, status = eng.align_many(refs, seqs, [c["cigar"] for c in cases])
    std, _, _ = eng.align_many(refs, seqs, [c["cigar"] for c in cases], standardize=True, collapse=True)
    bad += sum(o != c["out"] or s != c["std"] for o, s, c in zip(outs, std, cases))
    eng.get_np_info(refs[0])
    eng.close()
os.environ.pop("NPORE_TEAM", None)
# an SHR run beyond the 11-bit record field: second pass with the WIDE kernels + overflow list; packed (4-bit) read upload
from npore_b200.engine import PackedBatch, cigars_to_rle_batch
rng0 = np.random.default_rng(8)
a_, b_ = synth.make_reference(120, rng0, 0.0, "CGT"), synth.make_reference(120, rng0, 0.0, "CGT")
refl, seql = a_ + "A" * 12 + "A" * 2300 + b_, a_ + "A" * 12 + b_
cgl = "=" * 132 + "D" * 2300 + "=" * 120
eng = Realigner(S, NP)
o, _, st = eng.align_many([oracle.bases_to_int(refl)], [oracle.bases_to_int(seql)], [cgl])
bad += (o[0] != oracle.align(oracle.bases_to_int(refl), oracle.bases_to_int(seql), cgl, S, NP)) + int(st[0] != 0)
nib = np.zeros((len(seql) + 2) // 2 + 1, np.uint8)
codes16 = np.array([{"A": 1, "C": 2, "G": 4, "T": 8}[c] for c in seql], np.uint8)
for t, v in enumerate(codes16):          # the read starts at nibble 1 (an odd soft clip)
    k = t + 1
    nib[k >> 1] |= (v << 4) if k % 2 == 0 else v
words, off = cigars_to_rle_batch([cgl])
pk = PackedBatch.from_flat_shared_nib(oracle.bases_to_int(refl), np.zeros(1, np.int64), np.array([len(refl)], np.int32), nib, np.ones(1, np.int64),
                                      np.array([len(seql)], np.int32), words, off)
res = eng.align_packed(pk, 0, eng.new_result(pk, 0, pinned=False))
bad += res.ops_str(0) != o[0]
eng.close()
# a long item (several finish parts), expanded + run-length outputs, wide band (6-warp CTAs; r = 100: two-warp teams of <4,2>)
rng = np.random.default_rng(3)
cm = synth.call_length_model(NP)
ref, tr = synth.make_reference_with_tracts(60_000, rng)
rd = synth.make_reads(ref, 1, 45_000, rng, cm, tracts=tr)[0]
from npore_b200 import cig
for r, mb in ((30, 20000), (60, 5000), (100, 20000)):
    os.environ["NPORE_STD_LONG_MIN"] = "8" if r == 30 else "100000"      # segment-parallel standardisation on the first pass
    eng = Realigner(S, NP, max_b_rows=mb, r=r)
    ir, iq = oracle.bases_to_int(rd[9]), oracle.bases_to_int(rd[7])
    o, _, _ = eng.align_many([ir], [iq], [cig.expand_cigar(rd[5])], standardize=True)
    c, _, _ = eng.align_many([ir], [iq], [cig.expand_cigar(rd[5])], standardize=True, collapse=True)
    want = oracle.standardize(oracle.align(ir, iq, cig.expand_cigar(rd[5]), S, NP, max_b_rows=mb, r=r), ir, iq)
    bad += (o[0] != want) + (c[0] != oracle.collapse_cigar(want))
    eng.close()
# time-sliced teams: more chunks than resident teams, 24-step slices -- state save / restore, mailbox rebuild, team barriers
rs = synth.make_reads(ref, 100, 1200, rng, cm, tracts=tr)
os.environ["NPORE_RR_SLICE"] = "24"
for team, r, mb in (("2", 30, 48), ("4", 100, 120)):
    os.environ["NPORE_TEAM"] = team
    eng = Realigner(S, NP, max_b_rows=mb, r=r)
    irs, iqs, cgs = [oracle.bases_to_int(x[9]) for x in rs], [oracle.bases_to_int(x[7]) for x in rs], [cig.expand_cigar(x[5]) for x in rs]
    outs, _, _ = eng.align_many(irs, iqs, cgs)
    bad += sum(o != oracle.align(a_r, a_q, cg, S, NP, max_b_rows=mb, r=r) for o, a_r, a_q, cg in zip(outs, irs, iqs, cgs))
    eng.close()
os.environ.pop("NPORE_TEAM", None)
os.environ["NPORE_RR_SLICE"] = "40"
# confusion matrices
import pileup_oracle as po
from npore_b200 import confusion
contig = synth.make_reference(3000, rng, p_np=0.4)
reads = synth.make_aligned_reads(contig, 80, 400, rng)
want = po.confusion([po.Read(*r) for r in reads], contig, 100, 2900, oracle.get_np_info, oracle.bases_to_int)
got = confusion.calc_confusion_matrices_batch([("c", 100, 1500), ("c", 1500, 2900)], refs={"c": contig},
                                              reads={"c": confusion.AlignedReads([r[:5] for r in reads])})
want2 = None
for a, b in ((100, 1500), (1500, 2900)):
    one = po.confusion([po.Read(*r) for r in reads], contig, a, b, oracle.get_np_info, oracle.bases_to_int)
    want2 = one if want2 is None else tuple(x + y for x, y in zip(want2, one))
bad += sum(not np.array_equal(a, b) for a, b in zip(want2, got))
print("sanitize_run mismatches:", bad)
sys.exit(1 if bad else 0)
