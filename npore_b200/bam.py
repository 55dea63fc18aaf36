"""Batch equivalents of the reference's per-item work functions (/root/reference/src/bam.pyx:51-123), the seam the
two multiprocessing.Pool call sites are swapped for:

  realign.py:110-114        pool.imap_unordered(realign_read, read_data, chunksize=100)   ->  realign_reads(read_data)
  standardize_vcf.py:30-31  pool.map(realign_hap, hap_data)                               ->  realign_haps(hap_data)

Per item the GPU does align() + the CIGAR standardisation + collapse (bam.pyx:59-78); the host only packs inputs
and formats SAM text.  SAM records are emitted in INPUT order (the reference appends in completion order).
"""
import numpy as np

from . import cfg
from .aln import _engine, _report
from .cig import bases_to_int
from .engine import (NPORE_OUT_NO_EXPANDED, NPORE_OUT_RLE, NPORE_OUT_STANDARDIZE, PackedBatch, cigar_to_rle)
from .scheduler import iter_batches
from .confusion import calc_confusion_matrices, calc_confusion_matrices_batch, get_confusion_matrices, get_ranges  # noqa: F401  (bam.pyx:149-205, 351-510)


def sam_record(read_data, cigar_text):
    """bam.pyx:83, field for field (TLEN = reference span, RNEXT '*', PNEXT 0, tag HP:i)."""
    read_id, flag, ref_name, start, mapq, _, stop, seq, quals, _, hap = read_data
    return f"{read_id}\t{flag}\t{ref_name}\t{start + 1}\t{mapq}\t{cigar_text}\t*\t0\t{stop - start}\t{seq}\t{quals}\tHP:i:{hap}"


def _tables():
    if cfg.args.sub_scores is None or cfg.args.np_scores is None:
        raise RuntimeError("cfg.args.sub_scores / cfg.args.np_scores are not set (realign.py:92-93)")
    return cfg.args.sub_scores, cfg.args.np_scores


def realign_reads(read_data, write=True, max_batch_ops=64_000_000):
    """Realign an iterable of bam.pyx:34-47 tuples.  Returns the SAM record lines in input order; with write=True
    also appends them to f'{cfg.args.out_prefix}.sam' like realign_read does (bam.pyx:81-84)."""
    sub, npt = _tables()
    eng = _engine(sub, npt, 5, 1, 20000, 30)
    flags = NPORE_OUT_STANDARDIZE | NPORE_OUT_RLE | NPORE_OUT_NO_EXPANDED
    lines = []
    fh = open(f"{cfg.args.out_prefix}.sam", "a") if write else None
    try:
        for batch in iter_batches(read_data, lambda rd: len(rd[9]) + len(rd[7]), max_batch_ops):
            packed = PackedBatch.from_strings([rd[9] for rd in batch], [rd[7] for rd in batch], [rd[5] for rd in batch])
            res = eng.align_packed(packed, flags, eng.new_result(packed, flags, pinned=False))
            _report(res.status[:packed.n], "realign_read")
            new = [sam_record(rd, cg) for rd, cg in zip(batch, res.cigar_texts())]
            lines.extend(new)
            if fh:
                fh.write("\n".join(new) + "\n")
            with cfg.counter.get_lock():
                cfg.counter.value += len(batch)
    finally:
        if fh:
            fh.close()
    return lines


def realign_read(read_data):
    """bam.pyx:51-89: one read (a GPU batch of one); appends one SAM line, returns None."""
    realign_reads([read_data], write=True)


def realign_haps(hap_data, max_batch_ops=300_000_000):
    """bam.pyx:93-123 for a list of (contig, hap, seq, ref, expanded_cigar) -> same tuples with the standardised
    expanded CIGAR (over 'MID')."""
    hap_data = list(hap_data)
    from .engine import check_item_sizes
    check_item_sizes([len(h[3]) for h in hap_data], [len(h[2]) for h in hap_data], what="haplotype")     # per item, before any batch runs
    sub, npt = _tables()
    eng = _engine(sub, npt, 5, 1, 20000, 30)
    out = []
    for batch in iter_batches(hap_data, lambda h: len(h[2]) + len(h[3]), max_batch_ops):
        packed = PackedBatch.from_strings([h[3] for h in batch], [h[2] for h in batch], [h[4] for h in batch])
        res = eng.align_packed(packed, NPORE_OUT_STANDARDIZE, eng.new_result(packed, NPORE_OUT_STANDARDIZE, pinned=False))
        _report(res.status[:packed.n], "realign_hap")
        for k, (contig, hap, seq, ref, _) in enumerate(batch):
            out.append((contig, hap, seq, ref, res.ops_str(k)))
        with cfg.counter.get_lock():
            cfg.counter.value += len(batch)
    return out


def realign_hap(hap_data):
    """bam.pyx:93-123: one haplotype."""
    return realign_haps([hap_data])[0]
