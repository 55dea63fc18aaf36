"""Basecaller confusion matrices on the GPU -- host side of npore_confusion_batch (include/npore_b200.h).

Mirrors /root/reference/src/bam.pyx: get_ranges (149-164), calc_confusion_matrices (351-510) and
get_confusion_matrices (168-205).  The reference counts from the text of `samtools mpileup`; this path feeds the
alignments (position, CIGAR, bases, qualities) to the device, which derives the same events (npore_b200/csrc/confusion.cuh).
"""
import os

import numpy as np

from . import cfg
from .aln import _engine

SKIP_FLAGS = 0x4 | 0x100 | 0x200 | 0x400      # samtools mpileup --ff default: UNMAP, SECONDARY, QCFAIL, DUP
_REF_OPS = np.zeros(16, dtype=bool)
_REF_OPS[[0, 2, 3, 7, 8]] = True


class AlignedReads:
    """The reads of one contig as flat arrays, coordinate sorted (stable, i.e. BAM order for equal positions)."""

    def __init__(self, records):
        """records: iterable of (pos, cigar, seq, qual, flag) with cigar = uint32 BAM words or [(len, op char)],
        seq = str, qual = uint8 array / bytes / None (stored without qualities)."""
        recs = [r for r in records if not (r[4] & SKIP_FLAGS)]
        recs.sort(key=lambda r: r[0])
        n = len(recs)
        words = []
        for r in recs:
            c = r[1]
            if not isinstance(c, np.ndarray):
                c = np.array([(int(k) << 4) | cfg.cigar_dict[op] for k, op in c], dtype=np.uint32)
            words.append(c.astype(np.uint32, copy=False))
        self.pos = np.array([r[0] for r in recs], dtype=np.int64)
        self.cigar_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(w) for w in words], out=self.cigar_off[1:])
        self.cigar_rle = np.concatenate(words) if n and self.cigar_off[-1] else np.zeros(0, np.uint32)
        self.seq_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(r[2]) for r in recs], out=self.seq_off[1:])
        self.seq_ascii = np.frombuffer("".join(r[2] for r in recs).encode("latin-1"), dtype=np.uint8)
        q = [np.full(len(r[2]), 255, np.uint8) if r[3] is None else np.frombuffer(bytes(r[3]), dtype=np.uint8) for r in recs]
        self.qual = np.concatenate(q) if n and self.seq_off[-1] else np.zeros(0, np.uint8)
        span = np.where(_REF_OPS[self.cigar_rle & 15], self.cigar_rle >> 4, 0).astype(np.int64)
        csum = np.concatenate(([0], np.cumsum(span)))
        self.end = self.pos + (csum[self.cigar_off[1:]] - csum[self.cigar_off[:-1]])
        self.maxend = np.maximum.accumulate(self.end) if n else self.end

    def __len__(self):
        return len(self.pos)

    def overlapping(self, start, end):
        """Indices (ascending) of the reads with pos < end and end > start."""
        hi = int(np.searchsorted(self.pos, end, "left"))
        lo = int(np.searchsorted(self.maxend[:hi], start, "right"))
        return lo + np.flatnonzero(self.end[lo:hi] > start)


class PileupPack:
    """The arrays of one npore_pileup_batch."""

    def __init__(self, ranges, refs, reads, max_n, min_base_q=13):
        """ranges: [(ctg, start, end)]; refs: {ctg: str}; reads: {ctg: AlignedReads}."""
        ctgs = sorted({c for c, _, _ in ranges})
        base, off = {}, 0
        for c in ctgs:
            base[c] = off
            off += len(reads[c]) if c in reads else 0
        have = [reads[c] for c in ctgs if c in reads and len(reads[c])]
        cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)   # noqa: E731
        self.read_pos = cat([a.pos for a in have], np.int64)
        self.seq_ascii = cat([a.seq_ascii for a in have], np.uint8)
        self.qual = cat([a.qual for a in have], np.uint8)
        self.cigar_rle = cat([a.cigar_rle for a in have], np.uint32)
        self.seq_off = np.zeros(len(self.read_pos) + 1, np.int64)
        self.cigar_off = np.zeros(len(self.read_pos) + 1, np.int64)
        k = so = co = 0
        for a in have:
            self.seq_off[k + 1:k + 1 + len(a)] = a.seq_off[1:] + so
            self.cigar_off[k + 1:k + 1 + len(a)] = a.cigar_off[1:] + co
            k += len(a); so += int(a.seq_off[-1]); co += int(a.cigar_off[-1])
        self.range_start = np.array([s for _, s, _ in ranges], dtype=np.int64)
        self.range_end = np.array([e for _, _, e in ranges], dtype=np.int64)
        chunks, lists = [], []
        for c, s, e in ranges:
            contig = refs[c]
            if not (0 <= s < e <= len(contig)):
                raise ValueError(f"range {c}:{s}-{e} outside the contig (length {len(contig)})")
            chunks.append(contig[s:min(len(contig), e + 1 + max_n)])
            lists.append((reads[c].overlapping(s, e) + base[c]).astype(np.int32) if c in reads and len(reads[c]) else np.zeros(0, np.int32))
        self.ref_off = np.zeros(len(ranges) + 1, np.int64)
        np.cumsum([len(x) for x in chunks], out=self.ref_off[1:])
        self.ref_ascii = np.frombuffer("".join(chunks).encode("latin-1"), dtype=np.uint8)
        self.range_reads_off = np.zeros(len(ranges) + 1, np.int64)
        np.cumsum([len(x) for x in lists], out=self.range_reads_off[1:])
        self.range_reads = cat(lists, np.int32)
        self.min_base_q = min_base_q


def get_ranges(regions):
    """bam.pyx:149-164: cut every (contig, start, stop) into chunk_width windows."""
    w = int(cfg.args.chunk_width)
    return [(ctg, s, min(stop, s + w)) for ctg, start, stop in regions for s in range(start, stop, w)]


def _np_engine():
    z = np.zeros((int(cfg.args.max_n), int(cfg.args.max_l) + 1, int(cfg.args.max_l) + 1), np.float32)
    return _engine(np.zeros((5, 5), np.float32), z, 5, 1, 20000, 30)


def load_alignments(bam_fn):
    """{contig: AlignedReads} from a BAM file (npore_b200/bamio.py reader; no pysam / samtools)."""
    from .bamio import read_bam
    _, refs, recs = read_bam(bam_fn)
    per = {}
    for r in recs:
        if r["ref_id"] >= 0:
            per.setdefault(refs[r["ref_id"]][0], []).append((r["pos"], r["cigar"], r["seq"], r["qual"], r["flag"]))
    return {c: AlignedReads(v) for c, v in per.items()}


def calc_confusion_matrices_batch(ranges, refs=None, reads=None, min_base_q=13):
    """Sum of calc_confusion_matrices over `ranges` in one device call: (subs, nps, inss, dels), int64."""
    refs = cfg.args.refs if refs is None else refs
    if reads is None:
        reads = getattr(cfg.args, "_alignments", None)
        if reads is None:
            reads = cfg.args._alignments = load_alignments(cfg.args.bam)
    return _np_engine().confusion_batch(PileupPack(list(ranges), refs, reads, int(cfg.args.max_n), min_base_q))


def calc_confusion_matrices(range_tuple):
    """bam.pyx:351-510, same argument and return value (one (ctg, start, end) window)."""
    out = calc_confusion_matrices_batch([range_tuple])
    with cfg.counter.get_lock():
        cfg.counter.value += 1
    return out


def get_confusion_matrices():
    """bam.pyx:168-205: load the cached matrices from stats_dir, or count them from cfg.args.bam over
    get_ranges(cfg.args.regions) (one GPU batch instead of a process pool) and cache them."""
    d = cfg.args.stats_dir
    if not getattr(cfg.args, "recalc_cms", False):
        print("> loading confusion matrices")
        return tuple(np.load(os.path.join(d, f"{k}_cm.npy")) for k in ("subs", "nps", "inss", "dels"))
    print("> calculating confusion matrices")
    out = calc_confusion_matrices_batch(get_ranges(cfg.args.regions))
    os.makedirs(d, exist_ok=True)
    for k, m in zip(("subs", "nps", "inss", "dels"), out):
        np.save(os.path.join(d, f"{k}_cm"), m)
    return out
