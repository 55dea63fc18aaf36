"""CPU tests (-m "not gpu"): host logic, codecs, the C-ABI library's exported surface, and the N>1 sharding path
(world_size-2 gloo).  No GPU compute is attempted here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from npore_b200 import _lib, cig, scheduler, synth
from npore_b200.engine import PackedBatch, cigar_to_rle, rle_to_text

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """include/*.h <-> libnpore_b200.so: every declared entry point is exported (and listed in _lib.EXPORTS / IO_EXPORTS)."""
    L = ctypes.CDLL(_lib.LIB_PATH)
    for header, listed in (("npore_b200.h", _lib.EXPORTS), ("npore_bamio.h", _lib.IO_EXPORTS)):
        hdr = open(os.path.join(ROOT, "include", header)).read()
        declared = set(re.findall(r"\b(npore_[a-z_0-9]+)\s*\(", hdr))
        assert declared == set(listed), header
        for name in declared:
            assert hasattr(L, name), name
    assert sorted(os.listdir(os.path.join(ROOT, "include"))) == ["npore_b200.h", "npore_bamio.h"]
    assert b"sm_100a" in _lib.lib().npore_version()


def test_no_cpu_fallback_without_gpu(tables, have_gpu):
    """The product path must fail loudly when there is no CUDA device."""
    if have_gpu:
        pytest.skip("GPU present")
    from npore_b200.engine import NporeError, Realigner
    with pytest.raises(NporeError, match="no CUDA device"):
        Realigner(*tables)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "npore_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "ref_loader" not in src and "npore_oracle" not in src, fn


def test_cigar_codecs():
    assert cig.expand_cigar("1D3M2I") == "DMMMII"                      # cig.pyx:42-57 docstring
    assert cig.collapse_cigar("DMMMII") == "1D3M2I"                    # cig.pyx:13-38 docstring
    assert cig.collapse_cigar("DMMMII", return_groups=True) == [(1, "D"), (3, "M"), (2, "I")]
    assert cig.collapse_cigar("") == ""
    assert cig.bases_to_int("NACGT-x").tolist() == [0, 1, 2, 3, 4, 5, 0]
    assert cig.int_to_bases([1, 2, 3, 4, 0]) == "ACGTN"
    assert cig.seq_len("SXI=MD") == 5 and cig.ref_len("SXI=MD") == 4
    w = cigar_to_rle("3=2D1X4S2I")
    assert w.tolist() == [(3 << 4) | 7, (2 << 4) | 2, (1 << 4) | 8, (2 << 4) | 1]           # S dropped (bam.pyx:59)
    assert cigar_to_rle("===DDXII").tolist() == w.tolist()
    assert rle_to_text(w) == "3=2D1X2I"
    rng = np.random.default_rng(0)
    for _ in range(50):
        ops = "".join(rng.choice(list("=XIDM"), size=int(rng.integers(0, 200))))
        assert cig.expand_cigar(cig.collapse_cigar(ops)) == ops
        assert rle_to_text(cigar_to_rle(ops)) == cig.collapse_cigar(ops)


def test_packed_batch_layout():
    refs = [np.array([1, 2, 3], np.uint8), np.zeros(0, np.uint8), np.array([4, 4], np.uint8)]
    seqs = [np.array([1, 2], np.uint8), np.zeros(0, np.uint8), np.array([4, 4, 4], np.uint8)]
    rles = [cigar_to_rle("2=1D"), cigar_to_rle(""), cigar_to_rle("2=1I")]
    p = PackedBatch(refs, seqs, rles, pinned=False)
    assert p.ref_start.tolist() == [0, 3, 3] and p.ref_len.tolist() == [3, 0, 2] and p.ref_total == 5
    assert p.seq_start.tolist() == [0, 2, 2] and p.seq_total == 5
    assert p.cigar_off.tolist() == [0, 2, 2, 4] and p.total_ops == 10
    shared = np.arange(10, dtype=np.uint8) % 5
    q = PackedBatch(None, seqs, rles, shared_ref=shared, ref_ranges=[(0, 3), (3, 3), (5, 7)], pinned=False)
    assert q.ref_total == 10 and q.ref_len.tolist() == [3, 0, 2] and q.ref_start.tolist() == [0, 3, 5]


def test_synth_reads_are_consistent(tables):
    rng = np.random.default_rng(5)
    cm = synth.call_length_model(tables[1])
    ref, tr = synth.make_reference_with_tracts(20000, rng)
    for rd in synth.make_reads(ref, 5, 4000, rng, cm, tracts=tr):
        ops = cig.expand_cigar(rd[5])
        assert cig.ref_len(ops) == len(rd[9]) and cig.seq_len(ops) == len(rd[7])
        i = j = 0
        for op in ops:
            if op in "=X":
                assert (rd[9][j] == rd[7][i]) == (op == "=")
                i += 1; j += 1
            elif op == "I":
                i += 1
            else:
                j += 1


def test_scheduler_units():
    assert scheduler.n_chunks(0) == 0 and scheduler.n_chunks(19999) == 1 and scheduler.n_chunks(20000) == 2
    assert scheduler.n_cu(10000, 9950) == (19950 + 1) * 61
    batches = list(scheduler.iter_batches(range(10), lambda x: 4, 10))
    assert [len(b) for b in batches] == [2, 2, 2, 2, 2] and sum(batches, []) == list(range(10))
    sh = scheduler.shard_by_region(np.arange(100)[::-1], np.ones(100), 4)
    assert [len(s) for s in sh] == [25, 25, 25, 25]
    assert sorted(np.concatenate(sh).tolist()) == list(range(100))
    starts = np.arange(100)[::-1]
    assert all(starts[sh[g]].max() < starts[sh[g + 1]].min() for g in range(3))     # contiguous genomic regions


_DIST_SCRIPT = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from npore_b200 import scheduler
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(1)                      # same stream on every rank: the read set is global
starts = rng.integers(0, 1_000_000, size=400)
lens = rng.integers(5_000, 15_000, size=400)
loads = np.array([scheduler.n_cu(int(l), int(l)) for l in lens])
mine = scheduler.shard_by_region(starts, loads, world)[rank]
# every rank realigns only its shard (no collective on the data path); here: gather the bookkeeping to rank 0
t = torch.tensor([len(mine), int(loads[mine].sum()), int(starts[mine].min()), int(starts[mine].max())], dtype=torch.int64)
out = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(out, t)
if rank == 0:
    tot = sum(int(o[0]) for o in out)
    assert tot == 400, tot
    assert sum(int(o[1]) for o in out) == int(loads.sum())
    for a, b in zip(out[:-1], out[1:]):
        assert int(a[3]) <= int(b[2])                 # shards are ordered, non-overlapping regions
    mx = max(int(o[1]) for o in out); mean = loads.sum() / world
    assert mx < 1.05 * mean, (mx, mean)
    print("DIST_OK")
dist.destroy_process_group()
'''


def test_region_sharding_world_size_2(tmp_path):
    """N>1 path on CPU: torchrun, gloo, world_size 2 -- shards cover all reads once, are contiguous and balanced."""
    script = tmp_path / "dist_check.py"
    script.write_text(_DIST_SCRIPT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29541", str(script), ROOT], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "DIST_OK" in res.stdout


def test_batch_codecs_match_single_item_codecs():
    """The vectorised whole-batch packers / formatters used by realign_reads agree with the per-item functions."""
    from npore_b200.engine import bases_to_int_batch, cigars_to_rle_batch, rle_to_text_batch
    rng = np.random.default_rng(0)

    def rnd():
        k = int(rng.integers(0, 60))
        return "".join(f"{int(rng.integers(1, 30000 if rng.random() < 0.05 else 40))}{'MIDSH=X'[int(rng.integers(0, 7))]}" for _ in range(k))
    cigs = [rnd() for _ in range(300)] + ["", "5M", ""]
    w, off = cigars_to_rle_batch(cigs)
    assert all(np.array_equal(w[off[i]:off[i + 1]], cigar_to_rle(c)) for i, c in enumerate(cigs))
    assert rle_to_text_batch(w, off) == [rle_to_text(cigar_to_rle(c)) for c in cigs]
    ex = ["".join(rng.choice(list("=XIDM"), size=int(rng.integers(0, 300)))) for _ in range(100)] + ["", "I", "", ""]
    w2, off2 = cigars_to_rle_batch(ex)
    assert all(np.array_equal(w2[off2[i]:off2[i + 1]], cigar_to_rle(c)) for i, c in enumerate(ex))
    codes, lens = bases_to_int_batch(["ACGTN-x", "", "GGGTTT"])
    assert codes.tolist() == [1, 2, 3, 4, 0, 5, 0, 3, 3, 3, 4, 4, 4] and lens.tolist() == [7, 0, 6]
    p = PackedBatch.from_strings(["ACG", "", "TT"], ["AC", "", "TTT"], ["2=1D", "", "2=1I"])
    q = PackedBatch([cig.bases_to_int("ACG"), cig.bases_to_int(""), cig.bases_to_int("TT")],
                    [cig.bases_to_int("AC"), cig.bases_to_int(""), cig.bases_to_int("TTT")],
                    [cigar_to_rle("2=1D"), cigar_to_rle(""), cigar_to_rle("2=1I")], pinned=False)
    for f in ("ref_start", "ref_len", "seq_start", "seq_len", "cigar_off"):
        assert getattr(p, f).tolist() == getattr(q, f).tolist()
    assert p.ref_codes[:p.ref_total].tolist() == q.ref_codes[:q.ref_total].tolist()
    assert p.cigar_rle[:4].tolist() == q.cigar_rle[:4].tolist() and p.total_ops == q.total_ops


def test_oversize_item_is_reported_per_item_before_any_batch(monkeypatch):
    """one item with ref_len + seq_len >= 2^28 used to fail the whole npore_upload with NPORE_ERR_BAD_ARG,
    valid items included.  The size is checked per item, with the item named, before anything is sent to the device."""
    from npore_b200 import bam, engine
    engine.check_item_sizes([1 << 27], [(1 << 27) - 1])                       # exactly at the limit: accepted
    with pytest.raises(engine.NporeError, match=r"item 1: ref_len \+ seq_len = 268435456 exceeds"):
        engine.check_item_sizes([10, 1 << 27], [10, 1 << 27])
    monkeypatch.setattr(engine, "MAX_ITEM_OPS", 20)                           # realign_haps: checked before an engine exists
    haps = [("c", 1, "ACGT", "ACGT", "===="), ("c", 2, "ACGTACGTACGT", "ACGTACGTACGT", "=" * 12)]
    with pytest.raises(engine.NporeError, match=r"haplotype 1: ref_len \+ seq_len = 24 exceeds"):
        bam.realign_haps(haps)

