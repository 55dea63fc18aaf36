"""Boundary polish: the drop-in module names.  `PYTHONPATH=shim` makes `import aln / bam / cig / cfg` -- what
/root/reference/src/realign.py:11-13, bam.pyx:12-14 and test/align.py:9-12 import -- resolve to the GPU package, and cig carries the
reference's array helpers (cig.pyx:102-192, 234-256)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from npore_b200 import cig, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENV = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "shim") + os.pathsep + os.environ.get("PYTHONPATH", ""))


def test_shim_names_resolve_to_the_package():
    code = ("import aln, bam, cig, cfg, npore_b200.aln, npore_b200.cfg\n"
            "from aln import align, calc_score_matrices, dump, get_np_info, print_np_info, fix_matrix_properties\n"
            "from cig import expand_cigar, collapse_cigar, bases_to_int, cig_to_int, int_to_cig, push_indels_left, push_inss_thru_dels, same_cigar\n"
            "from bam import realign_read, realign_hap, get_confusion_matrices\n"
            "assert aln is npore_b200.aln and cfg is npore_b200.cfg\n"
            "import argparse; cfg.args = argparse.Namespace(max_n=3, max_l=50)\n"
            "assert npore_b200.cfg.args.max_n == 3\n"
            "print('SHIM_OK')\n")
    out = subprocess.run([sys.executable, "-c", code], env=ENV, capture_output=True, text=True, timeout=300)
    assert "SHIM_OK" in out.stdout, out.stdout + out.stderr


def test_cig_array_helpers_match_the_oracle(tables):
    """cig_to_int / int_to_cig / push_indels_left / push_inss_thru_dels / same_cigar composed as bam.pyx:65-78 composes them."""
    S, NP = tables
    cm = synth.call_length_model(NP)
    rng = np.random.default_rng(17)
    for _ in range(250):
        rf, sq, cg, _, _ = synth.fuzz_case(rng, cm)
        ir, iq = oracle.bases_to_int(rf), oracle.bases_to_int(sq)
        ex = cg.replace("M", "=")
        assert cig.standardize_cigar(ex, ir, iq) == oracle.standardize(ex, ir, iq)
    a = cig.cig_to_int("MMIDD=X")
    assert a.dtype == np.uint8 and a.tolist() == [0, 0, 1, 2, 2, 7, 8] and cig.int_to_cig(a) == "MMIDD=X"
    assert cig.same_cigar(a, a.copy()) and not cig.same_cigar(a, a[:-1])
    assert cig.int_to_cig(cig.push_inss_thru_dels(cig.cig_to_int("MDDIIM"))) == "MIIDDM"
    with pytest.raises(KeyError):
        cig.cig_to_int("MQ")


_ALIGN_PY = r'''
# the call sequence of /root/reference/test/align.py:9-60 (imports by the reference's module names; resolved by shim/)
import sys, json, argparse
import numpy as np
from aln import align, calc_score_matrices, dump
from cig import expand_cigar
import cfg as cfg
from bam import get_confusion_matrices
cfg.args = argparse.Namespace(recalc_cms=False, max_n=6, max_l=100, stats_dir=sys.argv[1])
subs, nps, inss, dels = get_confusion_matrices()
sub_scores, np_scores, ins_scores, del_scores = calc_score_matrices(subs, nps, inss, dels)
out = []
for ref, seq, cigar in json.load(open(sys.argv[2])):
    cigar = expand_cigar(cigar)
    int_ref = np.array([cfg.base_dict[c] for c in ref], dtype=np.uint8)
    int_seq = np.array([cfg.base_dict[c] for c in seq], dtype=np.uint8)
    new_cigar = align(int_ref, int_seq, cigar, sub_scores, np_scores, verbose=True, max_b_rows=20, r=10)
    dump(ref, seq, new_cigar)
    out.append(new_cigar)
print("RESULT " + json.dumps(out))
'''


@pytest.mark.gpu
def test_reference_test_script_flow_through_the_shim(golden, tmp_path):
    """test/align.py's flow (same imports, same calls incl. verbose=True, max_b_rows=20, r=10) with only PYTHONPATH=shim added:
    every KAT of the reference's own test file gives the reference's answer."""
    cm = np.load(os.path.join(ROOT, "tests", "golden", "cm_guppy5.npz"))
    stats = tmp_path / "stats"
    stats.mkdir()
    for k in ("subs", "nps", "inss", "dels"):
        np.save(str(stats / f"{k}_cm.npy"), cm[k])
    kats = golden("align_kats.json")
    (tmp_path / "cases.json").write_text(json.dumps([[k["ref"], k["seq"], k["cigar"]] for k in kats]))
    (tmp_path / "align_flow.py").write_text(_ALIGN_PY)
    out = subprocess.run([sys.executable, str(tmp_path / "align_flow.py"), str(stats), str(tmp_path / "cases.json")], env=ENV, capture_output=True,
                         text=True, timeout=600)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, out.stdout[-3000:] + out.stderr[-3000:]
    assert json.loads(line[0][7:]) == [k["small"]["out"] for k in kats]
    if os.path.exists("/root/reference/test/align.py"):          # the unmodified script itself, where the reference tree exists
        run = subprocess.run([sys.executable, "/root/reference/test/align.py", "--stats_dir", str(stats)], env=ENV, capture_output=True, text=True,
                             timeout=600, cwd="/root/reference/test")
        assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
