"""CPU tests (-m "not gpu"): pin the C oracle (oracle/npore_oracle.c) to the reference's own fixtures.

The golden vectors in tests/golden/ were produced by the UNMODIFIED reference (tests/golden/make_golden.py):
its one known-answer file test/data/npore_realigned.sam, the get_np_info docstring example (aln.pyx:182-194),
the inputs of test/align.py and test/get_np_info.py, test/test_std_vcf.vcf, and 400 seeded fuzz cases.
Where oracle/_ref is built (this container) the oracle is additionally fuzzed live against the reference."""
import numpy as np
import pytest

import oracle
from npore_b200 import synth


def test_golden_sam(tables, golden):
    """bam.pyx:51-89 realign_read on test/data/reads.sam + ref.fasta == test/data/npore_realigned.sam (all 10 records)."""
    S, NP = tables
    g = golden("golden_sam.json")
    assert len(g["reads"]) == 10
    for rd, want in zip(g["reads"], g["expected_sam"]):
        read_id, flag, ref_name, start, mapq, cigar, stop, seq, quals, ref, hap = rd
        cig = oracle.realign_cigar(ref, seq, cigar, S, NP)
        line = f"{read_id}\t{flag}\t{ref_name}\t{start + 1}\t{mapq}\t{cig}\t*\t0\t{stop - start}\t{seq}\t{quals}\tHP:i:{hap}"
        assert line == want


def test_np_info_kats(golden):
    for k in golden("np_info_kats.json"):
        info = oracle.get_np_info(oracle.bases_to_int(k["seq"]))
        assert info[:, 0, :].T.tolist() == k["L"]
        assert info[:, 1, :].T.tolist() == k["L_IDX"]


def test_np_info_docstring_example():
    """aln.pyx:182-194."""
    info = oracle.get_np_info(oracle.bases_to_int("ATATATATTTTTTAAAGCGCGC"))
    assert info[:, 0, 0].tolist() == [0, 0, 0, 0, 0, 0, 0, 6, 6, 6, 6, 6, 6, 3, 3, 3, 0, 0, 0, 0, 0, 0]
    assert info[:, 1, 0].tolist() == [0, 0, 0, 0, 0, 0, 0, 0, 1, 2, 3, 4, 5, 0, 1, 2, 0, 0, 0, 0, 0, 0]
    assert info[:, 0, 1].tolist() == [4, 3, 4, 3, 4, 3, 4, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 0, 3, 0, 3, 0]
    assert info[:, 1, 1].tolist() == [0, 0, 1, 1, 2, 2, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 2, 0]
    assert not info[:, :, 2:].any()


def test_align_kats(tables, golden):
    """test/align.py:20-39 cases at (max_b_rows=20, r=10) [test/align.py:59-60] and at align()'s defaults."""
    S, NP = tables
    for k in golden("align_kats.json"):
        ir, iq = oracle.bases_to_int(k["ref"]), oracle.bases_to_int(k["seq"])
        for tag in ("small", "default"):
            w = k[tag]
            out, sc, st = oracle.align(ir, iq, oracle.expand_cigar(k["cigar"]), S, NP, max_b_rows=w["max_b_rows"], r=w["r"],
                                       return_scores=True)
            assert out == w["out"] and st == 0
            assert np.array_equal(sc, np.array(w["scores"], dtype=np.float32))


def test_std_vcf_kats(tables, golden):
    """bam.pyx:93-123 realign_hap on the haplotypes of test/test_std_vcf.vcf."""
    S, NP = tables
    for k in golden("std_vcf_kats.json"):
        ir, iq = oracle.bases_to_int(k["ref"]), oracle.bases_to_int(k["seq"])
        out = oracle.standardize(oracle.align(ir, iq, oracle.expand_cigar(k["cigar"]), S, NP), ir, iq)
        assert out == k["out"]


def test_fuzz_golden(tables, golden):
    S, NP = tables
    for c in golden("fuzz.json.gz"):
        ir, iq = oracle.bases_to_int(c["ref"]), oracle.bases_to_int(c["seq"])
        out, sc, st = oracle.align(ir, iq, c["cigar"], S, NP, max_b_rows=c["max_b_rows"], r=c["r"], return_scores=True)
        assert out == c["out"] and st == 0
        assert np.array_equal(sc, np.array(c["scores"], dtype=np.float32))
        assert oracle.collapse_cigar(oracle.standardize(out, ir, iq)) == c["std"]


def test_len_tail_kats(tables, golden):
    """LEN on the last copies of a read tract (tests/golden/len_tail_kats.json, reference outputs + chunk scores)."""
    S, NP = tables
    for c in golden("len_tail_kats.json"):
        ir, iq = oracle.bases_to_int(c["ref"]), oracle.bases_to_int(c["seq"])
        got, sc, st = oracle.align(ir, iq, c["cigar"], S, NP, r=c["r"], max_n=c["max_n"], return_scores=True)
        assert got == c["out"] and st == 0 and np.array_equal(sc, np.array(c["scores"], dtype=np.float32))


def test_plan_matches_chunk_count(tables):
    """get_breaks (aln.pyx:344-358): number of chunks and the 'do not split a DI pair' shift."""
    assert oracle.plan("=" * 30, 30, 30, 20).tolist() == [0, 18, 38, 56, 60]      # breaks at 19k land on 'I' after 'D' -> shifted
    assert oracle.plan("I" * 5 + "D" * 7, 5, 7, 20000).tolist() == [0, 12]
    assert len(oracle.plan("", 0, 0, 20)) == 1


def test_live_against_reference(tables):
    """Differential fuzz against the compiled, unmodified reference (only where oracle/_ref is built)."""
    import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not built here")
    ref = ref_loader.load_reference()
    S, NP = tables
    cm = synth.call_length_model(NP)
    rng = np.random.default_rng(99)
    for _ in range(150):
        rf, sq, cg, r, mb = synth.fuzz_case(rng, cm)
        ir, iq = oracle.bases_to_int(rf), oracle.bases_to_int(sq)
        want, wsc = ref.aln_sc.align(ir, iq, cg, S, NP, 5, 1, mb, r)
        got, gsc, st = oracle.align(ir, iq, cg, S, NP, max_b_rows=mb, r=r, return_scores=True)
        assert got == want and st == 0
        assert np.array_equal(gsc, np.array(wsc, dtype=np.float32))
        if len(ir):
            assert np.array_equal(np.asarray(ref.aln.get_np_info(ir)), oracle.get_np_info(ir))


def test_pull_model_matches_oracle(tables):
    """oracle/pull_model.c (the gather-form / carried-BASE / chain-walker dataflow the GPU kernels implement) against
    the scatter-form oracle: op strings, bit-identical chunk scores, np_info."""
    import ctypes as C
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(oracle.__file__))
    subprocess.check_call(["make", "-C", here, "libpull_model.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(here, "libpull_model.so"))
    u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L.pm_align.restype = C.c_int64
    L.pm_align.argtypes = [u8, C.c_int, u8, C.c_int, C.c_char_p, C.c_int64, f32, f32, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                           C.c_int, C.c_int, C.c_char_p, C.c_int64, f32, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.pm_np_raw.argtypes = [u8, C.c_int, C.c_int, C.c_int, u8, np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")]
    S, NP = tables
    cm = synth.call_length_model(NP)
    rng = np.random.default_rng(31)
    pad = lambda a: a if a.size else np.zeros(1, np.uint8)  # noqa: E731
    for _ in range(200):
        rf, sq, cg, r, mb = synth.fuzz_case(rng, cm)
        ir, iq = oracle.bases_to_int(rf), oracle.bases_to_int(sq)
        want, wsc, wst = oracle.align(ir, iq, cg, S, NP, max_b_rows=mb, r=r, return_scores=True)
        cap = len(ir) + len(iq) + 8
        out = C.create_string_buffer(cap)
        sc = np.zeros(cap // max(1, mb - 1) + 4, np.float32)
        ns, st = C.c_int(0), C.c_int(0)
        cb = cg.encode()
        n = L.pm_align(pad(ir), len(ir), pad(iq), len(iq), cb, len(cb), S, NP, NP.shape[1], 6, 100, 5.0, 1.0, mb, r, out, cap, sc, len(sc),
                       C.byref(ns), C.byref(st))
        assert out.raw[:n].decode() == want and st.value == wst and np.array_equal(sc[:ns.value], wsc)
        if len(ir):
            raw = np.zeros((len(ir), 8), np.uint8); xs = np.zeros((len(ir), 8), np.int32)
            L.pm_np_raw(ir, len(ir), 6, 100, raw, xs)
            info = oracle.get_np_info(ir)
            assert np.array_equal(raw[:, :6] & 0x7f, info[:, 0, :]) and np.array_equal(xs[:, :6], info[:, 1, :])


def test_pull_model_v2_inf_ring_matches_oracle(tables):
    """oracle/pull_model.c:pm2_align -- the round-2 dataflow of csrc/forward.cuh (INF ring with physical slots, no source tests
    outside aliasing windows, 16-bit reciprocal, re-laid tables) -- against the scatter-form oracle, incl. band widths with 1-3
    spare slots and 6-12-base INDEL runs; the explicit-checks form (form 1) must agree too."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("pm2_fuzz", os.path.join(os.path.dirname(os.path.abspath(oracle.__file__)), "..", "tools", "pm2_fuzz.py"))
    pm2_fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pm2_fuzz)
    L = pm2_fuzz.load()
    S, NP = tables
    cm = synth.call_length_model(NP)
    rng = np.random.default_rng(2026)
    risky = 0
    for k in range(160):
        rf, sq, cg, r, mb = synth.fuzz_case(rng, cm)
        if k % 2:
            r = int(rng.choice([13, 14, 15, 29, 30, 31, 62, 63]))
        if k % 3 == 0:                                   # a long insertion into the input path (six equal ops in a row and more)
            pos = int(rng.integers(0, len(cg) + 1))
            ins = "".join(rng.choice(list("ACGT"), size=int(rng.integers(6, 12))))
            ro = sum(1 for c in cg[:pos] if c != "D")
            sq, cg = sq[:ro] + ins + sq[ro:], cg[:pos] + "I" * len(ins) + cg[pos:]
        ir, iq = oracle.bases_to_int(rf), oracle.bases_to_int(sq)
        want, wsc, wst = oracle.align(ir, iq, cg, S, NP, max_b_rows=mb, r=r, return_scores=True)
        for form in (0, 1):
            got, gsc, gst, nr = pm2_fuzz.pm2(L, ir, iq, cg, S, NP, mb, r, form)
            risky += nr
            assert got == want and gst == wst and np.array_equal(gsc, wsc), f"case {k} form {form} r={r} max_b_rows={mb}"
    assert risky > 0                                     # the checked path was exercised
