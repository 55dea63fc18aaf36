"""Genome-wide n-polymer regions from the device get_np_info (/root/reference/src/bed.py:56-145
without bedtools): the reference annotates the genome in `chunk_width` windows, lists (contig, start, start + n*L) for
every tract start, then pads (`slop`), sorts and merges per period n."""
import numpy as np

from . import cfg
from .aln import _engine
from .cig import bases_to_int


def _np_engine():
    t = cfg.args
    return _engine(np.zeros((5, 5), np.float32),
                   np.zeros((max(int(t.max_n), 1), int(t.max_l) + 1, int(t.max_l) + 1), np.float32), 5, 1, 20000, 30)


def get_np_regions_batch(regions, refs):
    """bed.py:56-76 for many (contig, start, stop) windows in one GPU launch.  refs: {contig: sequence}.
    Returns per window a list (per n) of (contig, start, stop) tuples."""
    eng = _np_engine()
    infos = eng.get_np_info_batch([bases_to_int(refs[c][a:b].upper()) for c, a, b in regions])
    out = []
    for (ctg, start, _), info in zip(regions, infos):
        per_n = []
        for n in range(1, info.shape[2] + 1):
            L, X = info[:, 0, n - 1], info[:, 1, n - 1]
            idx = np.flatnonzero((L != 0) & (X == 0))
            per_n.append([(ctg, int(start + p), int(start + p + n * L[p])) for p in idx])
        out.append(per_n)
    return out


def get_np_regions(region, refs=None):
    """bed.py:56-76, one window."""
    return get_np_regions_batch([region], refs if refs is not None else cfg.args.refs)[0]


def merge_regions(regions, slop=1):
    """What `bedtools merge` + sort do in bed.py:80-110: pad by slop, sort per contig, merge overlapping/adjacent."""
    by = {}
    for ctg, a, b in regions:
        by.setdefault(ctg, []).append((max(0, a - slop), b + slop))
    out = []
    for ctg in sorted(by, key=lambda c: (len(c), c)):
        cur = None
        for a, b in sorted(by[ctg]):
            if cur and a <= cur[1]:
                cur[1] = max(cur[1], b)
            else:
                if cur:
                    out.append((ctg, cur[0], cur[1]))
                cur = [a, b]
        if cur:
            out.append((ctg, cur[0], cur[1]))
    return out


def save_np_region_beds(np_regions, out_prefix, slop=1):
    """bed.py:80-145 (per-n BEDs + the union).  np_regions: list over windows of per-n lists."""
    max_n = len(np_regions[0]) if np_regions else 0
    union = []
    for n in range(1, max_n + 1):
        merged = merge_regions([r for win in np_regions for r in win[n - 1]], slop)
        union.extend(merged)
        with open(f"{out_prefix}_{n}.bed", "w") as fh:
            for ctg, a, b in merged:
                fh.write(f"{ctg}\t{a}\t{b}\n")
    with open(f"{out_prefix}_all.bed", "w") as fh:
        for ctg, a, b in merge_regions(union, 0):
            fh.write(f"{ctg}\t{a}\t{b}\n")
