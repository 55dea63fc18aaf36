"""Drop-in alias: `import bam` / `from bam import *` (as /root/reference/src/realign.py:11-13, bam.pyx:12-14 and test/align.py:9-12 do)
resolves to npore_b200.bam -- the same module object, so `cfg.args = parser.parse_args()` in a reference entry script is seen by
the GPU path.  Put this directory in front of the reference's src/ on PYTHONPATH."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import npore_b200.bam as _m  # noqa: E402

sys.modules[__name__] = _m
