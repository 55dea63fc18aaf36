"""Seeded synthetic workloads (pattern of /root/reference/test/generate_bam.py:34-101).

Used by bench.py and the tests.  Pure numpy/Python; nothing here is on the GPU hot path.
  * make_reference   : n-polymer-rich reference (D.1)
  * make_read        : ONT-like read with tract-length errors + sub/ins/del noise; CIGAR = true edit script (D.2)
  * make_reads       : 30x-style coverage over a reference (D.3)
  * fuzz_case        : small differential-fuzz cases (D.5)
"""
import numpy as np

BASES = "ACGT"


def make_reference(length: int, rng: np.random.Generator, p_np: float = 0.3, alphabet: str = BASES) -> str:
    parts, total = [], 0
    while total < length:
        if rng.random() < p_np:
            n = int(rng.integers(1, 7))
            unit = "".join(rng.choice(list(alphabet), size=n))
            copies = min(60, 3 + int(rng.geometric(0.25)))
            s = unit * copies
        else:
            s = "".join(rng.choice(list(alphabet), size=int(rng.integers(5, 41))))
        parts.append(s)
        total += len(s)
    return "".join(parts)[:length]


def _tracts(ref: str, max_n: int = 6):
    """Greedy left-to-right list of (start, n, copies) for >=3-copy period-n tracts (smallest n first)."""
    out, i, L = [], 0, len(ref)
    while i < L:
        hit = None
        for n in range(1, max_n + 1):
            if i + 3 * n > L:
                break
            unit = ref[i:i + n]
            c = 1
            while ref[i + c * n:i + (c + 1) * n] == unit:
                c += 1
            if c >= 3:
                hit = (i, n, c)
                break
        if hit:
            out.append(hit)
            i += hit[1] * hit[2]
        else:
            i += 1
    return out


def call_length_model(np_scores: np.ndarray):
    """P(called copies | n, ref copies) ~ exp(-np_scores[n-1, l, :]) -- derived from the learned table."""
    p = np.exp(-np_scores.astype(np.float64))
    p /= p.sum(axis=2, keepdims=True)
    return p


def make_read(ref: str, rng: np.random.Generator, call_model=None, p_ins=0.01, p_sub=0.015, p_del=0.015,
              alphabet: str = BASES):
    """Returns (seq, expanded_cigar over =XID) for a read covering the whole of `ref`."""
    tr = {s: (n, c) for s, n, c in _tracts(ref)}
    seq, cig, i, L = [], [], 0, len(ref)
    while i < L:
        t = tr.get(i)
        if t is not None and call_model is not None:
            n, c = t
            cc = min(c, call_model.shape[1] - 1)
            called = int(rng.choice(call_model.shape[2], p=call_model[n - 1, cc]))
            called = max(0, called + (c - cc))
            unit = ref[i:i + n]
            keep = min(c, called)
            seq.append(unit * keep)
            cig.append("=" * (keep * n))
            if called > c:
                seq.append(unit * (called - c))
                cig.append("I" * ((called - c) * n))
            elif called < c:
                cig.append("D" * ((c - called) * n))
            i += n * c
            continue
        u = rng.random()
        if u < p_ins:
            seq.append(alphabet[int(rng.integers(len(alphabet)))])
            cig.append("I")
        elif u < p_ins + p_del:
            cig.append("D")
            i += 1
        elif u < p_ins + p_del + p_sub:
            alt = [b for b in alphabet if b != ref[i]] or [ref[i]]
            seq.append(alt[int(rng.integers(len(alt)))])
            cig.append("X" if seq[-1] != ref[i] else "=")
            i += 1
        else:
            seq.append(ref[i])
            cig.append("=")
            i += 1
    return "".join(seq), "".join(cig)


def make_reads(reference: str, n_reads: int, read_len: int, rng: np.random.Generator, call_model=None):
    """List of read_data tuples shaped like bam.pyx:34-47:
    (read_id, flag, ref_name, start, mapq, cigarstring, stop, seq, quals, ref, hap)."""
    from .cig import collapse_cigar
    out, L = [], len(reference)
    starts = np.sort(rng.integers(0, max(1, L - read_len + 1), size=n_reads))
    for k, st in enumerate(starts):
        st = int(st)
        rs = reference[st:st + read_len]
        seq, cig = make_read(rs, rng, call_model)
        out.append((f"read{k}", 0, "ref", st, 60, collapse_cigar(cig), st + len(rs), seq, "*", rs, int(rng.integers(0, 3))))
    return out


def fuzz_case(rng: np.random.Generator, call_model=None):
    """One small (ref, seq, expanded_cigar, r, max_b_rows) differential-fuzz case (Appendix D.5)."""
    length = int(rng.integers(5, 700)) if rng.random() < 0.7 else int(rng.integers(5, 60))
    p_np = float(rng.choice([0.2, 0.6, 0.95]))
    alphabet = str(rng.choice(["ACGT", "AC", "ACGTN"]))
    ref = make_reference(length, rng, p_np, alphabet)
    if rng.random() < 0.05:   # very long tract: clamp quirk (max_l)
        k = int(rng.integers(0, len(ref)))
        ref = ref[:k] + str(rng.choice(list("ACGT"))) * int(rng.integers(101, 130)) + ref[k:]
    rate = float(rng.choice([0.0, 0.01, 0.05]))
    seq, cig = make_read(ref, rng, call_model if rng.random() < 0.8 else None, p_ins=rate, p_sub=rate * 1.5,
                         p_del=rate * 1.5, alphabet=alphabet)
    if rng.random() < 0.3:
        cig = cig.replace("=", "M").replace("X", "M")
    r = int(rng.choice([3, 5, 10, 30]))
    max_b_rows = int(rng.choice([12, 20, 37, 50, 200, 20000]))
    return ref, seq, cig, r, max_b_rows
