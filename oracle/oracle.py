"""ctypes front-end of the C oracle (oracle/npore_oracle.c)  -- TEST INFRASTRUCTURE ONLY.

Never imported by npore_b200/.  Function names mirror the reference's Python API
(aln.align, aln.get_np_info, cig.*; /root/reference/src/aln.pyx, cig.pyx, bam.pyx).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(HERE, "libnpore_oracle.so")
    src = os.path.join(HERE, "npore_oracle.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", HERE, "-B", "libnpore_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.npo_get_np_info.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _i32p]
        L.npo_align.restype = C.c_int64
        L.npo_align.argtypes = [_u8p, C.c_int, _u8p, C.c_int, C.c_char_p, C.c_int64, _f32p, _f32p, C.c_int,
                                C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                C.c_char_p, C.c_int64, _f32p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.npo_plan.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_int, _i32p, C.c_int]
        L.npo_standardize.restype = C.c_int64
        L.npo_standardize.argtypes = [C.c_char_p, C.c_int64, _u8p, _u8p, C.c_char_p]
        L.npo_collapse.restype = C.c_int64
        L.npo_collapse.argtypes = [C.c_char_p, C.c_int64, C.c_char_p]
        _LIB = L
    return _LIB


_CODE = np.zeros(256, dtype=np.uint8)
for _k, _c in enumerate("NACGT"):
    _CODE[ord(_c)] = _k
_CODE[ord("-")] = 5


def bases_to_int(seq: str) -> np.ndarray:
    """cig.pyx:212-229."""
    return _CODE[np.frombuffer(seq.encode(), dtype=np.uint8)].copy() if seq else np.zeros(0, np.uint8)


def expand_cigar(cigar: str) -> str:
    """cig.pyx:42-57."""
    out, count = [], 0
    for ch in cigar:
        if ch.isdigit():
            count = count * 10 + ord(ch) - 48
        else:
            out.append(ch * count)
            count = 0
    return "".join(out)


def _pad(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a if a.size else np.zeros(1, np.uint8)


def get_np_info(seq: np.ndarray, max_n=6, max_l=100) -> np.ndarray:
    """aln.pyx:179-251 -> int32 [len, 2, max_n]."""
    n = int(len(seq))
    out = np.zeros((max(n, 1), 2, max_n), dtype=np.int32)
    lib().npo_get_np_info(_pad(seq), n, max_n, max_l, out)
    return out[:n]


def align(full_ref, full_seq, cigar: str, sub_scores, np_scores, indel_start=5.0, indel_extend=1.0,
          max_b_rows=20000, r=30, max_n=6, max_l=100, return_scores=False):
    """aln.pyx:379-787.  Returns the expanded CIGAR (and per-chunk scores, status if asked)."""
    Lr, Ls = int(len(full_ref)), int(len(full_seq))
    sub = np.ascontiguousarray(sub_scores, dtype=np.float32)
    npt = np.ascontiguousarray(np_scores, dtype=np.float32)
    cap = Lr + Ls + 8
    out = C.create_string_buffer(cap)
    nchunk_cap = 2 + (Lr + Ls) // max(1, max_b_rows - 1) + 2
    scores = np.zeros(nchunk_cap, dtype=np.float32)
    nsc, st = C.c_int(0), C.c_int(0)
    cb = cigar.encode()
    n = lib().npo_align(_pad(full_ref), Lr, _pad(full_seq), Ls, cb, len(cb), sub, npt, int(npt.shape[1]),
                        max_n, max_l, indel_start, indel_extend, max_b_rows, r,
                        out, cap, scores, nchunk_cap, C.byref(nsc), C.byref(st))
    if n < 0:
        raise RuntimeError("oracle align failed")
    s = out.raw[:n].decode()
    if return_scores:
        return s, scores[:nsc.value].copy(), st.value
    return s


def plan(cigar: str, Ls: int, Lr: int, max_b_rows: int) -> np.ndarray:
    cb = cigar.encode()
    cap = 4 + (Ls + Lr) // max(1, max_b_rows - 1) + 2
    br = np.zeros(cap, dtype=np.int32)
    nb = lib().npo_plan(cb, len(cb), Ls, Lr, max_b_rows, br, cap)
    return br[:nb].copy()


def standardize(expanded: str, int_ref, int_seq) -> str:
    """bam.pyx:65-78: expanded {=,X,I,D} -> expanded standardised {M,I,D}."""
    cb = expanded.encode()
    out = C.create_string_buffer(len(cb) + 1)
    m = lib().npo_standardize(cb, len(cb), _pad(int_ref), _pad(int_seq), out)
    return out.raw[:m].decode()


def collapse_cigar(expanded: str) -> str:
    """cig.pyx:13-38."""
    cb = expanded.encode()
    out = C.create_string_buffer(len(cb) * 2 + 16)
    m = lib().npo_collapse(cb, len(cb), out)
    return out.raw[:m].decode()


def realign_cigar(ref: str, seq: str, cigar: str, sub_scores, np_scores, **kw) -> str:
    """The CIGAR column realign_read writes (bam.pyx:59-83): strip S/H, align, standardise, collapse."""
    ex = expand_cigar(cigar).replace("S", "").replace("H", "")
    ir, iq = bases_to_int(ref), bases_to_int(seq)
    return collapse_cigar(standardize(align(ir, iq, ex, sub_scores, np_scores, **kw), ir, iq))
