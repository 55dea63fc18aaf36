"""Drop-in for the hot-path functions of /root/reference/src/aln.pyx, executed on the GPU.

  align(full_ref, full_seq, cigar, sub_scores, np_scores, indel_start=5, indel_extend=1,
        max_b_rows=20000, r=30, verbose=0) -> str           aln.pyx:379-787   (batch of one on the B200)
  align_batch(...)                                          many align() calls in one GPU batch
  get_np_info(seq) -> int32 [len, 2, max_n]                 aln.pyx:179-251
  dump(ref, seq, cigar)                                     aln.pyx:791-...   (pretty printer, host)

Like the reference, max_n / max_l are read from cfg.args at call time.  Score-table construction
(calc_score_matrices, aln.pyx:62-96) is deliberately not re-implemented: its outputs are opaque inputs
of this path -- build them with the reference and pass the arrays in.
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
    """aln.pyx:379-382, same positional / keyword signature and return value (expanded CIGAR over '=XID')."""
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


def load_score_tables(path):
    """Tables saved with np.savez(path, sub_scores=..., np_scores=...) from the reference's calc_score_matrices."""
    t = np.load(path)
    return t["sub_scores"], t["np_scores"]
