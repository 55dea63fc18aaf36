"""Module-global configuration, mirroring /root/reference/src/cfg.py (same names) so the reference's entry scripts
(realign.py:119-121, standardize_vcf.py:88-90) can keep doing `cfg.args = parser.parse_args()`."""
import argparse
import threading
from collections import defaultdict

# globals read by the hot path at call time (aln.pyx:207-208, 436-437): max_n, max_l; tables live on
# args.sub_scores / args.np_scores (realign.py:92-93); out_prefix names the SAM (bam.pyx:82) and the log (aln.pyx:691)
args = argparse.Namespace(max_n=6, max_l=100, out_prefix="npore_out", sub_scores=None, np_scores=None, device=0,
                          chunk_width=100000, stats_dir="stats", recalc_cms=False, refs=None, bam=None, regions=None)


class _Counter:
    """Stand-in for the reference's mp.Value('i') (cfg.py:8): same .value / .get_lock() surface, thread lock."""

    def __init__(self):
        self.value = 0
        self._lock = threading.Lock()

    def get_lock(self):
        return self._lock


counter = _Counter()

bases = "NACGT"
symbols = "NACGT-"
nbases = len(bases)
base_dict = defaultdict(int, {c: i for i, c in enumerate("NACGT")})
base_dict.update({c.lower(): i for i, c in enumerate("NACGT")})
base_dict["-"] = 5

cigars = "MIDNSHP=XB"
cigar_dict = {c: i for i, c in enumerate(cigars)}

__version__ = "0.1.0-b200"
