"""Drop-in for the hot-path functions of /root/reference/src/aln.pyx, executed on the GPU.

  align(full_ref, full_seq, cigar, sub_scores, np_scores, indel_start=5, indel_extend=1,
        max_b_rows=20000, r=30, verbose=0) -> str           aln.pyx:379-787   (batch of one on the B200)
  align_batch(...)                                          many align() calls in one GPU batch
  get_np_info(seq) -> int32 [len, 2, max_n]                 aln.pyx:179-251
  dump(ref, seq, cigar)                                     aln.pyx:791-...   (pretty printer, host)

  calc_score_matrices(subs, nps, inss, dels, eps=0.01)     aln.pyx:62-96     (table construction, host, one-off)
  fix_matrix_properties(scores, delta=0.01)                 aln.pyx:11-58

Like the reference, max_n / max_l are read from cfg.args at call time.  The two table builders run once per
basecaller model on 60k numbers; they are host numpy and not part of the GPU path.
"""
import numpy as np

from . import cfg
from .engine import Realigner

_ENGINES = {}


def _engine(sub_scores, np_scores, indel_start, indel_extend, max_b_rows, r):
    sub = np.ascontiguousarray(sub_scores, dtype=np.float32)
    npt = np.ascontiguousarray(np_scores, dtype=np.float32)
    key = (sub.tobytes(), hash(npt.tobytes()), npt.shape, int(cfg.args.max_n), int(cfg.args.max_l), float(indel_start),
           float(indel_extend), int(max_b_rows), int(r), int(getattr(cfg.args, "device", 0) or 0))
    eng = _ENGINES.get(key)
    if eng is None:
        if len(_ENGINES) >= 8:
            _ENGINES.pop(next(iter(_ENGINES))).close()
        eng = Realigner(sub, npt, max_n=int(cfg.args.max_n), max_l=int(cfg.args.max_l), indel_start=float(indel_start),
                        indel_extend=float(indel_extend), max_b_rows=int(max_b_rows), r=int(r),
                        device=int(getattr(cfg.args, "device", 0) or 0))
        _ENGINES[key] = eng
    return eng


def _report(status, what="align"):
    """aln.pyx:689-716,737-739 print an ERROR line and return the partial CIGAR; same here."""
    names = {1: "row < 0", 2: "col < 0", 3: "run 0", 4: "unknown alignment matrix type", 16: "CIGAR inconsistent with sequence lengths"}
    for k, st in enumerate(np.atleast_1d(status)):
        if st:
            print(f"\nERROR: {names.get(int(st), st)} during traceback of item {k} ({what})")
            try:
                with open(f"{cfg.args.out_prefix}.log", "a+") as fh:
                    print(f"item: {k}, status: {int(st)}", file=fh)
            except OSError:
                pass


def align_batch(refs, seqs, cigars, sub_scores, np_scores, indel_start=5, indel_extend=1, max_b_rows=20000, r=30,
                return_scores=False):
    """Many align() calls as one GPU batch.  refs/seqs: uint8 code arrays; cigars: expanded or run-length text."""
    eng = _engine(sub_scores, np_scores, indel_start, indel_extend, max_b_rows, r)
    outs, scores, status = eng.align_many(refs, seqs, cigars)
    _report(status)
    return (outs, scores) if return_scores else outs


def align(full_ref, full_seq, cigar, sub_scores, np_scores, indel_start=5, indel_extend=1, max_b_rows=20000, r=30,
          verbose=0):
    """aln.pyx:379-382, same positional / keyword signature and return value (expanded CIGAR over '=XID').
    verbose: the reference prints its five (VAL, TYP, RUN) matrices (aln.pyx:744-785); they never exist here (2 bytes per cell
    reach memory), so verbose prints the per-chunk scores and the alignment instead."""
    if verbose:
        outs, scores = align_batch([np.asarray(full_ref, dtype=np.uint8)], [np.asarray(full_seq, dtype=np.uint8)], [cigar],
                                   sub_scores, np_scores, indel_start, indel_extend, max_b_rows, r, return_scores=True)
        print(f"align(verbose): {len(scores[0])} chunk(s), chunk scores {[float(x) for x in scores[0]]}; the DP matrices are not "
              "materialised on the GPU path (only the packed MAT record per cell)")
        print(f"align(verbose): CIGAR {outs[0]}")
        return outs[0]
    return align_batch([np.asarray(full_ref, dtype=np.uint8)], [np.asarray(full_seq, dtype=np.uint8)], [cigar],
                       sub_scores, np_scores, indel_start, indel_extend, max_b_rows, r)[0]


def get_np_info(seq):
    """aln.pyx:179-251 on device: int32 [len(seq), 2, max_n] with planes L (0) and L_IDX (1)."""
    t = cfg.args
    if t.sub_scores is not None and t.np_scores is not None:
        eng = _engine(t.sub_scores, t.np_scores, 5, 1, 20000, 30)
    else:   # np_info does not depend on the score tables
        eng = _engine(np.zeros((5, 5), np.float32), np.zeros((max(int(t.max_n), 1), int(t.max_l) + 1, int(t.max_l) + 1), np.float32),
                      5, 1, 20000, 30)
    return eng.get_np_info(np.asarray(seq, dtype=np.uint8))


def print_np_info(seq):
    """aln.pyx:(print_np_info): L / L_IDX rows per n."""
    info = get_np_info(seq)
    print("seq:   ", " ".join(f"{cfg.bases[b]:>2}" for b in seq))
    for n in range(info.shape[2]):
        print(f"N={n + 1} L: ", " ".join(f"{v:2}" for v in info[:, 0, n]))
        print("  L_IDX:", " ".join(f"{v:2}" for v in info[:, 1, n]))


def dump(ref, seq, cigar):
    """Pretty print an alignment (aln.pyx:791-...): three rows ref / ops / read."""
    ref_str, cig_str, seq_str = [], [], []
    i = j = 0
    for op in cigar:
        if op in "=XM":
            ref_str.append(ref[j]); seq_str.append(seq[i]); j += 1; i += 1
        elif op == "I":
            ref_str.append("-"); seq_str.append(seq[i]); i += 1
        elif op == "D":
            ref_str.append(ref[j]); seq_str.append("-"); j += 1
        cig_str.append(op)
    for k in range(0, len(cig_str), 80):
        print("REF: " + "".join(ref_str[k:k + 80]))
        print("     " + "".join(cig_str[k:k + 80]))
        print("SEQ: " + "".join(seq_str[k:k + 80]) + "\n")


def fix_matrix_properties(scores, delta=0.01):
    """aln.pyx:11-58.  In place, per period n, on the (ref_len, call_len) plane:
      1. rows 0..2 (no such tract) cost 20 for every call length >= 1; a correct call (the diagonal) costs 0;
      2. above the diagonal (insertions) a cell is at least delta worse than the cell below it and the cell to its left;
      3. below the diagonal, rows >= 4 (deletions): at least delta worse than the cell to its right and the cell above;
      4. rows >= 4: an off-diagonal cell is at least delta better than its upper-left neighbour (same INDEL, shorter tract).
    Sweeps 2-4 are recurrences evaluated in the reference's order with the array's own scalar type, so the result is
    the reference's to the bit."""
    size = scores.shape[1]
    for plane in scores:
        plane[0:3, 1:] = 20
        idx = np.arange(1, size)
        plane[idx, idx] = 0
        for col in range(1, size):
            for row in range(col - 1, -1, -1):
                plane[row, col] = max(plane[row, col], plane[row + 1, col] + delta, plane[row, col - 1] + delta)
        for row in range(4, size):
            for col in range(row - 1, -1, -1):
                plane[row, col] = max(plane[row, col], plane[row, col + 1] + delta, plane[row - 1, col] + delta)
        for row in range(4, size):
            for col in range(1, size):
                if row != col:
                    plane[row, col] = min(plane[row, col], plane[row - 1, col - 1] - delta)
    return scores


def _neg_log_frac(counts, totals, eps):
    return -np.log((np.asarray(counts, dtype=np.float64) + eps) / (np.asarray(totals, dtype=np.float64) + eps))


def calc_score_matrices(subs, nps, inss, dels, eps=0.01):
    """aln.pyx:62-96: confusion-matrix counts -> (sub_scores, np_scores, ins_scores, del_scores), all float32.
    score = -ln((count + eps) / (row total + eps)); only indices < max_l are filled (the last row / column of every
    table stays 0 before fix_matrix_properties, as in the reference); sub_scores row/column 0 (N) and diagonal are 0."""
    max_n, max_l = int(cfg.args.max_n), int(cfg.args.max_l)
    nps = np.asarray(nps)
    np_scores = np.zeros(nps.shape, dtype=np.float32)
    np_scores[:max_n, :max_l, :max_l] = _neg_log_frac(nps[:max_n, :max_l, :max_l], nps[:max_n, :max_l].sum(axis=2)[:, :, None], eps)
    np_scores = fix_matrix_properties(np_scores)

    subs = np.asarray(subs)
    sub_scores = np.zeros((cfg.nbases, cfg.nbases), dtype=np.float32)
    sub_scores[1:, 1:] = _neg_log_frac(subs[1:cfg.nbases, 1:cfg.nbases], subs[1:cfg.nbases].sum(axis=1)[:, None], eps)
    sub_scores[np.arange(cfg.nbases), np.arange(cfg.nbases)] = 0

    inss, dels = np.asarray(inss), np.asarray(dels)
    ins_scores = np.zeros(inss.shape, dtype=np.float32)
    ins_scores[:max_l] = _neg_log_frac(inss[:max_l], inss.sum(), eps)
    del_scores = np.zeros(dels.shape, dtype=np.float32)
    del_scores[:max_l] = _neg_log_frac(dels[:max_l], dels.sum(), eps)
    return sub_scores, np_scores, ins_scores, del_scores


def plot_np_score_matrices(nps, max_l=50):
    """aln.pyx:100-175 draws PNGs with matplotlib; plotting is outside this package's scope."""
    raise NotImplementedError("plotting is not part of npore_b200; use the reference's plot_np_score_matrices on the same arrays")


def load_score_tables(path):
    """Tables saved with np.savez(path, sub_scores=..., np_scores=...) from the reference's calc_score_matrices."""
    t = np.load(path)
    return t["sub_scores"], t["np_scores"]
