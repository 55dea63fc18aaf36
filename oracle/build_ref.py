"""Build the UNMODIFIED reference hot path (Cython) into oracle/_ref/  -- TEST INFRASTRUCTURE ONLY.

The reference (TimD1/nPoRe) implements the path in three Cython modules
(/root/reference/src/aln.pyx, cig.pyx, bam.pyx; recipe /root/reference/setup.py:4-13).
This script cythonizes them *where they lie* (no source is copied into the repository;
generated C and the built .so files go to the git-ignored oracle/_ref/), plus
  * cfg.py        -> compiled to cfg.*.so so no reference .py has to travel,
  * aln_sc        -> a build-time patched copy of aln.pyx (generated into oracle/_ref/gen/,
                     git-ignored) that additionally returns the per-chunk DP scores
                     (SURVEY.md section 8(c)): the reference's align() returns the CIGAR only,
  * stats/*.npy   -> copies of guppy5_stats (inputs of calc_score_matrices) for the
                     `bench.py --impl reference` arm on the GPU box.
Nothing under oracle/ is ever imported by the product package (npore_b200/).
Only tests/, __graft_entry__.smoke()/build() and bench.py's cpu_baseline / --impl reference
legs may use it.
"""
import os
import shutil
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("NPORE_REFERENCE", "/root/reference")


def _patched_aln_source(src: str) -> str:
    """Return aln.pyx text with three score-reporting edits (applied by anchor text)."""
    a1 = "    # iterate over b matrix in chunks set by breakpoints\n"
    a2 = "        b_col = a_to_b_col(a_row, a_col, inss, dels, r)\n        run = 0\n        path = []\n"
    a3 = "    return full_aln\n"
    assert src.count(a1) == 1 and src.count(a2) == 1 and src.count(a3) == 1, "reference aln.pyx changed"
    src = src.replace(a1, "    chunk_scores = []\n" + a1)
    src = src.replace(a2, a2 + "        chunk_scores.append(float(matrix[MAT, b_row, b_col, VAL]))\n")
    src = src.replace(a3, "    return full_aln, chunk_scores\n")
    return src


def ref_is_built() -> bool:
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(OUT, m + suffix)) for m in ("aln", "cig", "bam", "cfg", "aln_sc"))


def build(force: bool = False) -> bool:
    """Build oracle/_ref. Returns True if the reference modules are available afterwards."""
    if ref_is_built() and not force:
        return True
    src_dir = os.path.join(REF, "src")
    if not os.path.isdir(src_dir):
        return ref_is_built()
    from setuptools import Extension
    from setuptools.dist import Distribution
    from Cython.Build import cythonize

    os.makedirs(os.path.join(OUT, "gen"), exist_ok=True)
    with open(os.path.join(src_dir, "aln.pyx")) as fh:
        patched = _patched_aln_source(fh.read())
    with open(os.path.join(OUT, "gen", "aln_sc.pyx"), "w") as fh:
        fh.write(patched)

    exts = [
        Extension("aln", [os.path.join(src_dir, "aln.pyx")]),
        Extension("cig", [os.path.join(src_dir, "cig.pyx")]),
        Extension("bam", [os.path.join(src_dir, "bam.pyx")]),
        Extension("cfg", [os.path.join(src_dir, "cfg.py")]),
        Extension("aln_sc", [os.path.join(OUT, "gen", "aln_sc.pyx")]),
    ]
    build_dir = os.path.join(OUT, "build")
    ext_modules = cythonize(exts, language_level="3str", build_dir=build_dir, quiet=True)
    dist = Distribution({"name": "npore_ref", "ext_modules": ext_modules})
    cmd = dist.get_command_obj("build_ext")
    cmd.build_lib = OUT
    cmd.build_temp = build_dir
    cmd.ensure_finalized()
    cmd.run()

    stats_src = os.path.join(REF, "guppy5_stats")
    stats_dst = os.path.join(OUT, "stats")
    os.makedirs(stats_dst, exist_ok=True)
    for fn in os.listdir(stats_src):
        if fn.endswith(".npy"):
            shutil.copy(os.path.join(stats_src, fn), os.path.join(stats_dst, fn))
    return ref_is_built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built" if ok else "oracle/_ref NOT available (no /root/reference)")
    sys.exit(0 if ok else 1)
