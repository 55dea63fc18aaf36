"""The multi-GPU row: region sharding of one coordinate-sorted read set, independent per-GPU pipelines,
ordered host gather.  CPU tests cover the cut and the shared-buffer gather with two gloo ranks; the gpu test runs the product
path bamio.realign_bam(devices=[...]) and demands a SAM byte-identical to the single-GPU one."""
import os
import subprocess
import sys

import numpy as np
import pytest

from npore_b200 import scheduler

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cut_balanced_is_contiguous_and_balanced():
    rng = np.random.default_rng(5)
    loads = rng.integers(500_000, 1_500_000, size=5000)
    for world in (1, 2, 3, 4, 8):
        own = scheduler.cut_balanced(loads, world)
        assert own.min() == 0 and own.max() == world - 1 and (np.diff(own) >= 0).all()
        per = np.array([loads[own == g].sum() for g in range(world)], dtype=np.float64)
        assert abs(per - loads.sum() / world).max() <= loads.max()
    # cutting a list that arrives in pieces (one contig at a time) gives the same owners
    a = scheduler.cut_balanced(loads, 4)
    b = np.concatenate([scheduler.cut_balanced(loads[:1234], 4, 0.0, float(loads.sum())),
                        scheduler.cut_balanced(loads[1234:], 4, float(loads[:1234].sum()), float(loads.sum()))])
    assert np.array_equal(a, b)
    assert len(scheduler.cut_balanced(np.zeros(0), 4)) == 0


_GATHER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
path, n_tiles, cap = {path!r}, 6, 1000
if rank == 0:
    with open(path, "wb") as fh: fh.truncate(n_tiles * cap * 4)
dist.barrier()
buf = torch.from_file(path, shared=True, size=n_tiles * cap, dtype=torch.int32).numpy()
mine = range(n_tiles * rank // world, n_tiles * (rank + 1) // world)        # contiguous region of this rank (bench.py run_c3)
for t in mine:
    rng = np.random.default_rng(100 + t)
    buf[t * cap:(t + 1) * cap] = rng.integers(0, 1 << 30, size=cap)          # "the device-to-host copy IS the gather"
dist.barrier()
if rank == 0:
    want = np.concatenate([np.random.default_rng(100 + t).integers(0, 1 << 30, size=cap) for t in range(n_tiles)])
    assert np.array_equal(buf, want.astype(np.int32)), "gathered buffer is not the region-ordered concatenation"
    os.remove(path)
    print("GATHER_OK")
dist.destroy_process_group()
"""


def test_shared_buffer_gather_world_size_2(tmp_path):
    """Two gloo ranks write their regions' result slots into ONE shared host buffer; rank 0 reads the region-ordered whole."""
    script = tmp_path / "g.py"
    script.write_text(_GATHER.format(root=ROOT, path=f"/dev/shm/npore_test_gather_{os.getpid()}"))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", str(29600 + os.getpid() % 300), str(script)], capture_output=True, text=True, timeout=300)
    assert "GATHER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.gpu
def test_sharded_realign_bam_is_byte_identical(tables, tmp_path):
    """One coordinate-sorted BAM (two contigs), realigned on 1 device and as 2 / 3 / G region shards (every visible GPU; a device
    may carry several shards): the merged SAM files are byte-identical, shard loads are balanced, every shard did work."""
    import re
    import torch
    from npore_b200 import bamio, cfg, synth
    S, NP = tables
    cfg.args.sub_scores, cfg.args.np_scores = S, NP
    cfg.args.max_n, cfg.args.max_l = 6, 100
    rng = np.random.default_rng(31337)
    cm = synth.call_length_model(NP)
    contigs, recs = [], []
    for ci, (name, length, n_reads) in enumerate((("chrA", 120_000, 260), ("chrB", 60_000, 110))):
        ref, tr = synth.make_reference_with_tracts(length, rng)
        contigs.append((name, ref))
        for rd in synth.make_reads(ref, n_reads, 3000, rng, cm, tracts=tr):
            recs.append({"name": f"{name}_{rd[0]}", "flag": 0, "ref_id": ci, "pos": rd[3], "mapq": 60, "seq": rd[7], "qual": bytes([25] * len(rd[7])),
                         "cigar": [(int(a), b) for a, b in re.findall(r"(\d+)(\D)", rd[5])], "tags": {"HP": rd[10]}})
    bam = str(tmp_path / "in.bam")
    bamio.write_bam(bam, "@HD\tVN:1.6\tSO:coordinate\n", [(n, len(s)) for n, s in contigs], recs)
    fa = dict(contigs)
    n1 = bamio.realign_bam(bam, fa, out_prefix=str(tmp_path / "one"), argv=["t"], max_batch_ops=900_000)
    one = open(str(tmp_path / "one.sam"), "rb").read()
    assert n1 == len(recs) and one.count(b"\n") == len(recs) + 4
    ngpu = torch.cuda.device_count()
    for devices in ([0, 0], [k % ngpu for k in range(3)], list(range(max(ngpu, 2)))[:8] if ngpu > 1 else [0, 0, 0, 0]):
        devices = [d % ngpu for d in devices]
        tm = {}
        n = bamio.realign_bam(bam, fa, out_prefix=str(tmp_path / "many"), argv=["t"], max_batch_ops=900_000, devices=devices, timings=tm)
        many = open(str(tmp_path / "many.sam"), "rb").read()
        assert n == len(recs) and many == one, f"devices={devices}: merged SAM differs from the single-GPU SAM"
        sh = tm["shards"]
        assert len(sh) == len(devices) and sum(s["reads"] for s in sh) == len(recs) and all(s["reads"] > 0 for s in sh)
        loads = np.array([s["cell_updates"] for s in sh], dtype=np.float64)
        assert loads.max() / loads.mean() < 1.15
        assert not [f for f in os.listdir(tmp_path) if ".part" in f]
