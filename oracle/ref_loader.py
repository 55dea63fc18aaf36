"""Import the compiled, unmodified reference modules from oracle/_ref  -- TEST INFRASTRUCTURE ONLY.

Usage:  ref = load_reference();  ref.aln.align(...), ref.bam.realign_read(...), ref.aln_sc.align(...)
The reference imports pysam / matplotlib / Bio / vcf / util at module import time
(/root/reference/src/aln.pyx:4, cig.pyx:5-8, bam.pyx:7-15); none of them is touched by the
hot path, so empty stub modules are injected before import.
"""
import argparse
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_loaded = None


class Reference:
    pass


def available() -> bool:
    sys.path.insert(0, HERE)
    try:
        import build_ref
        return build_ref.ref_is_built()
    finally:
        sys.path.remove(HERE)


def load_reference(max_n: int = 6, max_l: int = 100, out_prefix: str = "/tmp/npore_ref_out"):
    """Returns an object with .aln .cig .bam .cfg .aln_sc, cfg.args initialised like realign.py:119-121."""
    global _loaded
    if _loaded is not None:
        _loaded.cfg.args.max_n = max_n
        _loaded.cfg.args.max_l = max_l
        _loaded.cfg.args.out_prefix = out_prefix
        return _loaded
    if not available():
        raise RuntimeError("oracle/_ref is not built (run python oracle/build_ref.py where /root/reference exists)")
    for m in ("pysam", "matplotlib", "matplotlib.pyplot", "Bio", "Bio.SeqIO", "vcf", "util"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["Bio"].SeqIO = sys.modules["Bio.SeqIO"]
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import cfg, cig, aln, bam, aln_sc  # noqa: E401  (compiled reference modules)
    ref = Reference()
    ref.cfg, ref.cig, ref.aln, ref.bam, ref.aln_sc = cfg, cig, aln, bam, aln_sc
    cfg.args = argparse.Namespace(
        max_n=max_n, max_l=max_l, stats_dir=os.path.join(REF_DIR, "stats"),
        recalc_cms=False, out_prefix=out_prefix, max_reads=0)
    _loaded = ref
    return ref


def reference_tables(ref=None):
    """sub_scores[5,5], np_scores[6,101,101] exactly as realign.py:87-93 builds them."""
    ref = ref or load_reference()
    subs, nps, inss, dels = ref.bam.get_confusion_matrices()
    sub_scores, np_scores, _, _ = ref.aln.calc_score_matrices(subs, nps, inss, dels)
    ref.cfg.args.sub_scores = sub_scores
    ref.cfg.args.np_scores = np_scores
    return sub_scores, np_scores
