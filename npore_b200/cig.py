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
