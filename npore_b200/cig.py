"""CIGAR text codecs -- host-side mirror of /root/reference/src/cig.pyx:13-57, 212-241 (same names).

These are the string<->array helpers either side of the GPU path; the device kernels work on
one-byte-per-op arrays.  Vectorised with numpy (the reference's are per-character Python loops).
"""
import re

import numpy as np

_CODE = np.zeros(256, dtype=np.uint8)
for _k, _c in enumerate("NACGT"):
    _CODE[ord(_c)] = _k
_CODE[ord("-")] = 5
_RLE = re.compile(r"(\d+)(\D)")


def bases_to_int(seq: str) -> np.ndarray:
    """cig.pyx:212-229: N,A,C,G,T,- -> 0..5; anything else -> 0."""
    if not seq:
        return np.zeros(0, dtype=np.uint8)
    return _CODE[np.frombuffer(seq.encode("latin-1"), dtype=np.uint8)]


def int_to_bases(int_seq) -> str:
    """cig.pyx:231-232."""
    return "".join("NACGT"[i] for i in int_seq)


def expand_cigar(cigar: str) -> str:
    """cig.pyx:42-57: '1D3M2I' -> 'DMMMII'."""
    return "".join(op * int(n) for n, op in _RLE.findall(cigar))


def collapse_cigar(extended_cigar: str, return_groups: bool = False):
    """cig.pyx:13-38: 'DMMMII' -> '1D3M2I'."""
    if not extended_cigar:
        return [] if return_groups else ""
    a = np.frombuffer(extended_cigar.encode("latin-1"), dtype=np.uint8)
    cut = np.flatnonzero(a[1:] != a[:-1]) + 1
    starts = np.concatenate(([0], cut))
    lens = np.diff(np.concatenate((starts, [len(a)])))
    groups = [(int(n), chr(a[s])) for n, s in zip(lens, starts)]
    if return_groups:
        return groups
    return "".join(f"{n}{op}" for n, op in groups)


def seq_len(cigar: str) -> int:
    """cig.pyx:196-201."""
    return sum(1 for op in cigar if op in "SXI=M")


def ref_len(cigar: str) -> int:
    """cig.pyx:203-208."""
    return sum(1 for op in cigar if op in "XD=M")


# ---------------------------------------------------------------------------------------------------------------------
# Array forms of the CIGAR (cig.pyx:60-72, 91-98, 234-256) and the two standardisation sweeps (cig.pyx:102-192).  The GPU path
# runs the standardisation of bam.pyx:65-78 on the device in the run-length domain (csrc/finish.cuh); these host versions keep the
# reference's per-op API for callers that hold uint8 op arrays (`from cig import *` in realign.py:12 / bam.pyx:13).
class Cigar:
    """pysam's cigartuples op codes (cig.pyx:60-72)."""
    M, I, D, N, S, H, P, E, X, B = range(10)      # noqa: E741


_OP_CODE = np.full(256, 255, dtype=np.uint8)
for _k, _c in enumerate("MIDNSHP=XB"):            # cfg.cigars / cfg.cigar_dict (cfg.py:27-32)
    _OP_CODE[ord(_c)] = _k
_OP_CHAR = np.frombuffer(b"MIDNSHP=XB", dtype=np.uint8)


def cig_to_int(cig: str) -> np.ndarray:
    """cig.pyx:234-238: 'MMID' -> uint8 [0, 0, 1, 2] (KeyError on a character outside 'MIDNSHP=XB', like cfg.cigar_dict)."""
    a = _OP_CODE[np.frombuffer(cig.encode("latin-1"), dtype=np.uint8)] if cig else np.zeros(0, np.uint8)
    if len(a) and a.max() == 255:
        raise KeyError(cig[int(np.argmax(a == 255))])
    return a.copy()


def int_to_cig(int_cig) -> str:
    """cig.pyx:240-241."""
    return _OP_CHAR[np.asarray(int_cig, dtype=np.uint8)].tobytes().decode("latin-1")


def extend_pysam_cigar(ops, counts) -> str:
    """cig.pyx:91-98: (['D','M','I'] as op codes, [1, 3, 2]) -> 'DMMMII'."""
    return "".join(int(n) * "MIDNSHP=XB"[op] for n, op in zip(counts, ops))


def same_cigar(cig1, cig2) -> bool:
    """cig.pyx:245-256."""
    return len(cig1) == len(cig2) and bool(np.array_equal(np.asarray(cig1), np.asarray(cig2)))


def push_indels_left(cigar, seq, nshifts_buf=None, shiftlen_buf=None, push_op=Cigar.D):
    """cig.pyx:102-159, in place on a uint8 op array; returns it.  Left to right, every maximal run of `push_op` is rotated left
    through as many preceding M / '=' ops as keep the sequence unchanged: a shift by one more position is allowed while the base
    leaving the run on the right equals the base entering it on the left (seq[p-1] == seq[p-1+len]).  `seq` is the sequence the
    run's ops consume (the reference for D, the read for I); its pointer advances over M / X / = ops and over the runs of push_op
    already passed, NOT over ops of the other indel type.  The two scratch buffers of the reference's signature are accepted and
    ignored (the rotation is done with one temporary)."""
    cg = np.asarray(cigar)
    n, sq = len(cg), np.asarray(seq)
    c = s = 0
    while c < n:
        op = int(cg[c])
        if op != push_op:
            c += 1
            if op in (Cigar.M, Cigar.X, Cigar.E):
                s += 1
            continue
        ln = 1
        while c + ln < n and cg[c + ln] == push_op:
            ln += 1
        k = 0
        while c - k > 0 and s - k > 0 and sq[s - k - 1] == sq[s - k - 1 + ln] and cg[c - k - 1] in (Cigar.M, Cigar.E):
            k += 1
        if k:
            moved = cg[c - k:c].copy()
            cg[c - k:c - k + ln] = push_op
            cg[c - k + ln:c + ln] = moved
        c += ln
        s += ln
    return cigar


def push_inss_thru_dels(cigar):
    """cig.pyx:164-192, in place; returns the array.  Every D immediately followed by I: the maximal block D..D I..I around that
    boundary is rewritten as I..I D..D, and the scan continues with the next position."""
    cg = np.asarray(cigar)
    n = len(cg)
    for i in range(n - 1):
        if cg[i] == Cigar.D and cg[i + 1] == Cigar.I:
            lo = i
            while lo - 1 >= 0 and cg[lo - 1] == Cigar.D:
                lo -= 1
            hi = i + 1
            while hi + 1 < n and cg[hi + 1] == Cigar.I:
                hi += 1
            n_ins = hi - i
            cg[lo:lo + n_ins] = Cigar.I
            cg[lo + n_ins:hi + 1] = Cigar.D
    return cigar


def standardize_cigar(cigar: str, int_ref, int_seq) -> str:
    """The standardisation block of bam.pyx:65-78 (= 105-118) on the host, built from the functions above: expanded '=XID' in,
    expanded 'MID' out.  (The loop body of the reference runs exactly once: `old_cig = int_cig[:]` is a view of the same
    buffer, so same_cigar is always true.)"""
    int_cig = cig_to_int(cigar.replace("X", "M").replace("=", "M"))
    push_indels_left(int_cig, int_ref, None, None, Cigar.D)
    push_inss_thru_dels(int_cig)
    push_indels_left(int_cig, int_seq, None, None, Cigar.I)
    push_inss_thru_dels(int_cig)
    return int_to_cig(int_cig).replace("ID", "M")
