"""Fresh seeded differential fuzz (not the committed fixtures): CUDA path vs the C oracle -- op strings, chunk scores,
standardised + collapsed CIGARs -- over random band radii / window sizes / time-slice lengths, small cases and 1-6 kb reads,
with the long-item standardisation forced on half of the groups.  usage: python tools/gpu_fuzz_live.py [n_groups] [seed] [n_production_reads] [n_long_items]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402
from npore_b200 import cig, synth  # noqa: E402
from npore_b200.engine import Realigner  # noqa: E402


def main():
    n_groups = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
    t = np.load(os.path.join(ROOT, "tests/golden/tables.npz")); S, NP = t["sub_scores"], t["np_scores"]
    cm = synth.call_length_model(NP)
    rng = np.random.default_rng(seed)
    ref, tr = synth.make_reference_with_tracts(300_000, rng)
    bad = total = chunks = n_info = 0
    fails = []
    t0 = time.time()
    for g in range(n_groups):
        r = int(rng.choice([3, 5, 8, 10, 15, 16, 20, 30, 31, 32, 45, 60, 64, 90, 127]))
        mb = int(rng.choice([12, 20, 37, 64, 100, 257, 1000, 5000, 20000, 65000]))
        os.environ["NPORE_RR_SLICE"] = str(int(rng.choice([7, 40, 512, 100000])))
        os.environ["NPORE_STD_LONG_MIN"] = "1" if g % 2 else "4096"
        max_n = int(rng.choice([6, 6, 6, 3, 1]))
        # every few groups: other gap penalties, and a smaller max_l with a correspondingly smaller table (np_scores[:, :l+1, :l+1])
        gopen, gext = (5.0, 1.0) if g % 5 else (float(rng.choice([2.0, 3.5, 8.0])), float(rng.choice([0.25, 0.5, 2.0])))
        max_l = 100 if g % 7 else int(rng.choice([20, 50, 99]))
        NPg = np.ascontiguousarray(NP[:, :max_l + 1, :max_l + 1])
        cases = []
        for _ in range(int(rng.integers(10, 60))):
            rf, sq, cg, _, _ = synth.fuzz_case(rng, cm)
            cases.append((rf, sq, cg))
        for _ in range(int(rng.integers(5, 30))):       # tract-centred: unit of 1-6 bases, 3-140 copies, copy-number change + noise
            alpha = list("ACGT") if rng.random() < 0.8 else list("ACGTN")
            unit = "".join(rng.choice(list("ACGT"), size=int(rng.integers(1, 7))))
            copies = int(rng.integers(3, max(4, 140 // len(unit))))
            pre = "".join(rng.choice(alpha, size=int(rng.integers(0, 30)))); suf = "".join(rng.choice(alpha, size=int(rng.integers(0, 30))))
            rf = pre + unit * copies + suf
            sq0 = pre + unit * int(rng.integers(max(0, copies - 12), copies + 12)) + suf
            sq, _ = synth.make_read(sq0, rng, None, p_ins=0.03, p_sub=0.03, p_del=0.03, alphabet="".join(alpha)) if sq0 else ("", "")
            m = min(len(rf), len(sq))
            style = rng.random()
            if style < 0.5:
                cg = "M" * m + "D" * (len(rf) - m) + "I" * (len(sq) - m)
            elif style < 0.75:
                cg = "D" * (len(rf) - m) + "I" * (len(sq) - m) + "M" * m
            else:
                h = m // 2
                cg = "=" * h + "I" * (len(sq) - m) + "D" * (len(rf) - m) + "X" * (m - h)
            cases.append((rf, sq, cg))
        if mb >= 257:
            for rd in synth.make_reads(ref, int(rng.integers(1, 5)), int(rng.integers(1000, 6000)), rng, cm, tracts=tr):
                cases.append((rd[9], rd[7], cig.expand_cigar(rd[5])))
        eng = Realigner(S, NPg, max_b_rows=mb, r=r, max_n=max_n, max_l=max_l, indel_start=gopen, indel_extend=gext)
        refs = [oracle.bases_to_int(c[0]) for c in cases]; seqs = [oracle.bases_to_int(c[1]) for c in cases]
        outs, scores, status = eng.align_many(refs, seqs, [c[2] for c in cases])
        std, _, _ = eng.align_many(refs, seqs, [c[2] for c in cases], standardize=True, collapse=True)
        raw_rle, _, _ = eng.align_many(refs, seqs, [c[2] for c in cases], collapse=True)            # '=XID' run-length, not standardised
        exp_std, _, _ = eng.align_many(refs, seqs, [c[2] for c in cases], standardize=True)          # expanded 'MID'
        for k, c in enumerate(cases):
            want, wsc, wst = oracle.align(refs[k], seqs[k], c[2], S, NPg, gopen, gext, max_b_rows=mb, r=r, max_n=max_n, max_l=max_l, return_scores=True)
            ws = oracle.collapse_cigar(oracle.standardize(want, refs[k], seqs[k]))
            ok = outs[k] == want and status[k] == wst and np.array_equal(scores[k], np.asarray(wsc, np.float32)) and std[k] == ws
            ok = ok and raw_rle[k] == oracle.collapse_cigar(want) and oracle.collapse_cigar(exp_std[k]) == ws
            total += 1; chunks += len(wsc); bad += (not ok)
            if not ok:
                what = [n for n, f in (("ops", outs[k] != want), ("status", status[k] != wst), ("scores", not np.array_equal(scores[k], np.asarray(wsc, np.float32))),
                                       ("std", std[k] != ws)) if f]
                print(f"MISMATCH group {g} case {k}: r={r} mb={mb} max_n={max_n} max_l={max_l} gap={gopen}/{gext} slice={os.environ['NPORE_RR_SLICE']} "
                      f"long_min={os.environ['NPORE_STD_LONG_MIN']} len={len(c[0])}/{len(c[1])} differs: {what} status={status[k]}/{wst}", flush=True)
                fails.append({"ref": c[0], "seq": c[1], "cigar": c[2], "r": r, "mb": mb, "max_n": max_n, "max_l": max_l, "gap": [gopen, gext], "slice": os.environ["NPORE_RR_SLICE"],
                              "long_min": os.environ["NPORE_STD_LONG_MIN"], "got_ops": outs[k], "want_ops": want, "got_std": std[k], "want_std": ws,
                              "got_scores": [float(x) for x in scores[k]], "want_scores": [float(x) for x in wsc], "n_in_group": len(cases), "k": k})
        # get_np_info of every sequence of the group (one CTA each) against the oracle
        both = [x for x in refs + seqs if len(x)]
        for x, got_info in zip(both, eng.get_np_info_batch(both)):
            n_info += 1
            if not np.array_equal(got_info, oracle.get_np_info(x, max_n=max_n, max_l=max_l)):
                bad += 1
                print(f"MISMATCH np_info group {g}: max_n={max_n} max_l={max_l} len={len(x)}", flush=True)
        eng.close()
    if fails:
        import json
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(fails, open(os.path.join(ROOT, "gpurun_out", f"fuzz_fail_{seed}.json"), "w"))
    n_reads = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    if n_reads:        # production-style reads (learned copy-number error model) at align()'s defaults, one GPU batch
        for k in ("NPORE_RR_SLICE", "NPORE_STD_LONG_MIN"):
            os.environ.pop(k, None)
        reads = synth.make_reads(ref, n_reads, int(rng.integers(2000, 6000)), rng, cm, tracts=tr)
        eng = Realigner(S, NP)
        refs = [oracle.bases_to_int(r[9]) for r in reads]; seqs = [oracle.bases_to_int(r[7]) for r in reads]
        cgs = [cig.expand_cigar(r[5]) for r in reads]
        outs, scores, status = eng.align_many(refs, seqs, cgs)
        std, _, _ = eng.align_many(refs, seqs, cgs, standardize=True, collapse=True)
        for k in range(n_reads):
            want, wsc, wst = oracle.align(refs[k], seqs[k], cgs[k], S, NP, return_scores=True)
            ok = outs[k] == want and status[k] == wst and np.array_equal(scores[k], np.asarray(wsc, np.float32)) and \
                std[k] == oracle.collapse_cigar(oracle.standardize(want, refs[k], seqs[k]))
            total += 1; chunks += len(wsc); bad += (not ok)
            if not ok:
                print(f"MISMATCH production read {k}", flush=True)
        eng.close()
    n_long = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    if n_long:         # long items: several plan / finish parts per item, many chunks per item, long-item standardisation
        big_ref, big_tr = synth.make_reference_with_tracts(1_500_000, rng)
        for mb, rr_, hap in ((5000, 30, False), (20000, 30, False), (50000, 10, False), (20000, 30, True)):
            os.environ["NPORE_STD_LONG_MIN"] = "4096"
            items = []
            for _ in range(n_long):
                L = int(rng.integers(20_000, 130_000))
                if hap:
                    st = int(rng.integers(0, len(big_ref) - L))
                    sub = big_ref[st:st + L]
                    keep = np.array([(a - st, b, c) for a, b, c in big_tr if st <= a and a + b * c <= st + L and rng.random() < 0.5]).reshape(-1, 3)
                    sq, cg = synth.make_read(sub, rng, cm, p_ins=0.0, p_sub=0.0005, p_del=0.0, tracts=keep)
                    items.append((sub, sq, cg))
                else:
                    rd = synth.make_reads(big_ref, 1, L, rng, cm, tracts=big_tr)[0]
                    items.append((rd[9], rd[7], cig.expand_cigar(rd[5])))
            eng = Realigner(S, NP, max_b_rows=mb, r=rr_)
            refs = [oracle.bases_to_int(c[0]) for c in items]; seqs = [oracle.bases_to_int(c[1]) for c in items]
            outs, scores, status = eng.align_many(refs, seqs, [c[2] for c in items])
            std, _, _ = eng.align_many(refs, seqs, [c[2] for c in items], standardize=True, collapse=True)
            exp, _, _ = eng.align_many(refs, seqs, [c[2] for c in items], standardize=True)
            for k, c in enumerate(items):
                want, wsc, wst = oracle.align(refs[k], seqs[k], c[2], S, NP, max_b_rows=mb, r=rr_, return_scores=True)
                wstd = oracle.standardize(want, refs[k], seqs[k])
                ok = outs[k] == want and status[k] == wst and np.array_equal(scores[k], np.asarray(wsc, np.float32)) and \
                    std[k] == oracle.collapse_cigar(wstd) and exp[k] == wstd
                total += 1; chunks += len(wsc); bad += (not ok)
                if not ok:
                    what = [n for n, f in (("ops", outs[k] != want), ("status", status[k] != wst), ("scores", not np.array_equal(scores[k], np.asarray(wsc, np.float32))),
                                           ("rle", std[k] != oracle.collapse_cigar(wstd)), ("expanded", exp[k] != wstd)) if f]
                    print(f"MISMATCH long item {k}: mb={mb} r={rr_} hap={hap} len={len(c[0])}/{len(c[1])} differs: {what} status {status[k]}/{wst}", flush=True)
                    import json
                    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                    json.dump({"ref": c[0], "seq": c[1], "cigar": c[2], "mb": mb, "r": rr_, "got_ops": outs[k], "want_ops": want, "got_std": std[k],
                               "got_exp": exp[k], "want_std": wstd, "got_scores": [float(x) for x in scores[k]], "want_scores": [float(x) for x in wsc]},
                              open(os.path.join(ROOT, "gpurun_out", f"long_fail_{seed}_{k}.json"), "w"))
            eng.close()
    print(f"live fuzz seed {seed}: {n_groups} groups, {total} cases, {chunks} chunks, {n_info} np_info arrays, {bad} mismatches, {time.time() - t0:.0f} s")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
