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


def make_reference_with_tracts(length: int, rng: np.random.Generator, p_np: float = 0.3, alphabet: str = BASES):
    """Like make_reference, also returning the emitted tracts as an int array [T,3] = (start, n, copies)."""
    parts, tr, total = [], [], 0
    letters = np.array(list(alphabet))
    while total < length:
        if rng.random() < p_np:
            n = int(rng.integers(1, 7))
            unit = "".join(letters[rng.integers(0, len(letters), size=n)])
            copies = min(60, 3 + int(rng.geometric(0.25)))
            s = unit * copies
            if total + len(s) <= length:
                tr.append((total, n, copies))
        else:
            s = "".join(letters[rng.integers(0, len(letters), size=int(rng.integers(5, 41)))])
        parts.append(s)
        total += len(s)
    return "".join(parts)[:length], np.array(tr, dtype=np.int64).reshape(-1, 3)


def call_length_cdf(call_model):
    return np.cumsum(call_model, axis=2)


def make_read(ref: str, rng: np.random.Generator, call_model=None, p_ins=0.01, p_sub=0.015, p_del=0.015,
              alphabet: str = BASES, tracts=None):
    """ONT-like read over the whole of `ref` (Appendix D.2), vectorised.  Returns (seq, expanded CIGAR in =XID).
    tracts: int array [T,3] of (start, n, copies) inside ref (non-overlapping); scanned from ref if None.
    Tract copy-number errors are drawn from call_model[n-1, copies] and placed at the tract end; every other
    base gets independent ins-before / sub / del noise."""
    L = len(ref)
    if L == 0:
        return "", ""
    if tracts is None:
        tracts = np.array(_tracts(ref), dtype=np.int64).reshape(-1, 3)
    rb = np.frombuffer(ref.encode(), dtype=np.uint8)
    letters = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    op = np.full(L, ord("="), dtype=np.uint8)
    base = rb.copy()
    n_ins = np.zeros(L, dtype=np.int64)          # bases inserted AFTER ref position p
    ins_src = np.full(L, -1, dtype=np.int64)     # >=0: inserted bases copy ref[ins_src + k % ins_n]; -1: random
    ins_n = np.ones(L, dtype=np.int64)
    noisy = np.ones(L, dtype=bool)
    if call_model is not None and len(tracts):
        st, n, c = tracts[:, 0], tracts[:, 1], tracts[:, 2]
        cc = np.minimum(c, call_model.shape[1] - 1)
        cdf = np.cumsum(call_model[n - 1, cc], axis=1)
        called = (cdf < rng.random(len(st))[:, None]).sum(axis=1)
        called = np.maximum(0, np.minimum(called, call_model.shape[2] - 1) + (c - cc))
        for s0, n0, c0, k0 in zip(st.tolist(), n.tolist(), c.tolist(), called.tolist()):
            e0 = s0 + n0 * c0
            noisy[s0:e0] = False
            if k0 < c0:
                op[e0 - (c0 - k0) * n0:e0] = ord("D")
            elif k0 > c0:
                n_ins[e0 - 1] = (k0 - c0) * n0
                ins_src[e0 - 1] = s0
                ins_n[e0 - 1] = n0
    m = int(noisy.sum())
    if m and (p_ins or p_sub or p_del):
        u = rng.random(m)
        idx = np.flatnonzero(noisy)
        dele = u < p_del
        sub = (u >= p_del) & (u < p_del + p_sub)
        op[idx[dele]] = ord("D")
        si = idx[sub]
        if len(si) and len(letters) > 1:
            cur = np.searchsorted(np.sort(letters), base[si])
            srt = np.sort(letters)
            base[si] = srt[(cur + rng.integers(1, len(letters), size=len(si))) % len(letters)]
            op[si] = np.where(base[si] != rb[si], ord("X"), ord("="))
        ins = idx[rng.random(m) < p_ins]
        ins = ins[n_ins[ins] == 0]
        n_ins[ins] = 1
    # assemble: for every ref position its op (and base unless D), then its inserted bases
    counts = np.empty(2 * L, dtype=np.int64)
    counts[0::2] = 1
    counts[1::2] = n_ins
    cig2 = np.empty(2 * L, dtype=np.uint8)
    cig2[0::2] = op
    cig2[1::2] = ord("I")
    cig = np.repeat(cig2, counts)
    total = int(counts.sum())
    owner = np.repeat(np.arange(2 * L), counts)              # which slot produced each output op
    pos = owner >> 1
    is_ins = (owner & 1).astype(bool)
    first = np.cumsum(counts) - counts                        # start offset of each slot
    k = np.arange(total) - first[owner]
    seqb = base[pos].copy()
    src = ins_src[pos]
    tract_ins = is_ins & (src >= 0)
    seqb[tract_ins] = rb[src[tract_ins] + k[tract_ins] % ins_n[pos[tract_ins]]]
    rnd = is_ins & (src < 0)
    seqb[rnd] = letters[rng.integers(0, len(letters), size=int(rnd.sum()))]
    keep = cig != ord("D")
    return seqb[keep].tobytes().decode(), cig.tobytes().decode()


def make_reads(reference: str, n_reads: int, read_len: int, rng: np.random.Generator, call_model=None, tracts=None):
    """List of read_data tuples shaped like bam.pyx:34-47:
    (read_id, flag, ref_name, start, mapq, cigarstring, stop, seq, quals, ref, hap).  Sorted by start."""
    from .cig import collapse_cigar
    out, L = [], len(reference)
    starts = np.sort(rng.integers(0, max(1, L - read_len + 1), size=n_reads))
    if tracts is None:
        tracts = np.array(_tracts(reference), dtype=np.int64).reshape(-1, 3)
    tends = tracts[:, 0] + tracts[:, 1] * tracts[:, 2] if len(tracts) else np.zeros(0, np.int64)
    for k, st in enumerate(starts.tolist()):
        en = min(L, st + read_len)
        rs = reference[st:en]
        if len(tracts):
            a = int(np.searchsorted(tracts[:, 0], st, side="left"))
            b = int(np.searchsorted(tends, en, side="right"))
            tr = tracts[a:b].copy()
            tr[:, 0] -= st
        else:
            tr = tracts
        seq, cig = make_read(rs, rng, call_model, tracts=tr)
        out.append((f"read{k}", 0, "ref", st, 60, collapse_cigar(cig), en, seq, "*", rs, int(rng.integers(0, 3))))
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


def make_aligned_reads(ref: str, n_reads: int, read_len: int, rng: np.random.Generator, max_n: int = 6,
                       p_tract: float = 0.3, with_clips: bool = True):
    """Coordinate-sorted alignments against `ref` for the confusion-matrix path (bam.pyx:351-510): every read is
    (pos, [(len, op)], seq, qual bytes | None, flag, mapq).  Unlike make_read, copy-number errors sit at the START of a
    tract (where a left-aligning mapper reports them and where the reference counts them), and the script mixes in the
    cases the pileup text is sensitive to: D next to I, I next to D, N read bases, low base qualities, reads without
    qualities, reverse / secondary reads, soft clips and uncovered stretches."""
    L = len(ref)
    letters = "ACGT"
    tract_at = {}
    for s0, n0, c0 in _tracts(ref, max_n):
        tract_at.setdefault(s0, (n0, c0))
    reads = []
    for _ in range(n_reads):
        span = int(min(L, max(1, read_len + rng.integers(-read_len // 4, read_len // 4 + 1))))
        pos = int(rng.integers(0, L - span + 1))
        ops, seq = [], []

        def emit(op, n, bases=""):
            if n <= 0:
                return
            if ops and ops[-1][1] == op:
                ops[-1][0] += n
            else:
                ops.append([n, op])
            seq.append(bases)

        eqx = rng.random() < 0.3                   # '=' / 'X' instead of 'M'
        if with_clips and rng.random() < 0.15:
            emit("H", int(rng.integers(1, 6)))
        if with_clips and rng.random() < 0.3:
            k = int(rng.integers(1, 6))
            emit("S", k, "".join(rng.choice(list(letters), size=k)))
            if rng.random() < 0.3:
                emit("I", 2, "".join(rng.choice(list(letters), size=2)))
        p, end = pos, pos + span
        while p < end:
            u = rng.random()
            b = ref[p]
            if u < 0.02:
                b = letters[(letters.find(b) + int(rng.integers(1, 4))) % 4] if b in letters else "A"
            elif u < 0.023:
                b = "N"
            elif u < 0.0235:
                b = "R"                            # IUPAC code: the reference's parser abandons the pileup line there
            emit(("=" if b == ref[p] else "X") if eqx else "M", 1, b)
            p += 1
            if p >= end:
                break
            tr = tract_at.get(p)
            u = rng.random()
            if tr is not None and u < p_tract:
                n0, c0 = tr
                k = int(rng.integers(1, 4))
                if rng.random() < 0.5:
                    d = min(k * n0, end - p - 1)
                    emit("D", d)
                    p += d
                else:
                    unit = ref[p:p + n0]
                    ins = unit * k
                    if rng.random() < 0.15:      # not a clean copy
                        ins = ins[:-1] + letters[(letters.find(ins[-1]) + 1) % 4]
                    emit("I", len(ins), ins)
            elif u < 0.012:
                d = min(int(rng.integers(1, 4)), end - p - 1)
                emit("N" if rng.random() < 0.05 else "D", d)
                p += d
                if rng.random() < 0.25:           # D directly followed by I
                    k = int(rng.integers(1, 3))
                    emit("I", k, "".join(rng.choice(list(letters), size=k)))
            elif u < 0.024:
                k = int(rng.integers(1, 4))
                emit("I", k, "".join(rng.choice(list(letters), size=k)))
                if rng.random() < 0.25 and end - p > 3:   # I directly followed by D
                    emit("D", 1)
                    p += 1
        if with_clips and rng.random() < 0.3:
            k = int(rng.integers(1, 6))
            emit("S", k, "".join(rng.choice(list(letters), size=k)))
        s = "".join(seq)
        qual = None if rng.random() < 0.1 else bytes(rng.integers(3, 41, size=len(s)).astype(np.uint8))
        flag = (16 if rng.random() < 0.5 else 0) | (256 if rng.random() < 0.05 else 0) | (1024 if rng.random() < 0.02 else 0)
        reads.append((pos, [(n, op) for n, op in ops], s, qual, flag, int(rng.integers(0, 61))))
    reads.sort(key=lambda r: r[0])
    return reads
