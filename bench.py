#!/usr/bin/env python
"""bench.py -- throughput of the realignment hot path (nPoRe align(): aln.pyx:379-787 + per-read glue) on B200.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                    (the reference's own CPU implementation, host cores)

A "step" = one pass of the hot path over one batch of synthetic reads.  Workload at N=1 = BASELINE.json
configs[1] ("C2"): seeded 1 Mb n-polymer-rich reference, 3,000 ONT-like 10 kb reads (30x), align() defaults
(r=30, max_b_rows=20000) + CIGAR standardisation + collapse, i.e. what realign_read does per read.  For N>1 every
rank gets its own C2-sized region shard (different seed): weak scaling, no collective on the data path.

Printed JSON line (rank 0): metric GCUPS (cell updates / s); `value` = kernels only, inputs
resident in HBM, timed with CUDA events on the launching stream; `e2e` = through the C-ABI call
npore_align_batch with pinned HOST buffers (H2D + kernels + D2H inside the timed region).  `bam_to_sam` (N=1 only,
informational) = a BAM file of the first 1,000 reads in, a realigned SAM file out (bamio.realign_bam).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OPS_PER_CU = 28
BYTES_PER_CU = 2         # packed (TYP,RUN) traceback record
NCU_DRAM_BYTES_PER_LAUNCH = 8.659e9   # forward_kernel<2> on the C2 batch: 7.968 GB written + 0.691 GB read (ncu, profiles/r01_final_*)


def load_tables():
    t = np.load(os.path.join(ROOT, "tests", "golden", "tables.npz"))
    return t["sub_scores"], t["np_scores"]


def make_workload(seed, ref_len, n_reads, read_len, np_scores):
    from npore_b200 import synth
    rng = np.random.default_rng(seed)
    cm = synth.call_length_model(np_scores)
    ref, tracts = synth.make_reference_with_tracts(ref_len, rng)
    reads = synth.make_reads(ref, n_reads, read_len, rng, cm, tracts=tracts)
    return ref, reads


def pack_reads(reads, pinned=True):
    from npore_b200.cig import bases_to_int
    from npore_b200.engine import PackedBatch, cigar_to_rle
    refs = [bases_to_int(r[9]) for r in reads]
    seqs = [bases_to_int(r[7]) for r in reads]
    rles = [cigar_to_rle(r[5]) for r in reads]
    return PackedBatch(refs, seqs, rles, pinned=pinned)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        """Started BEFORE the warm-up steps (nvidia-smi takes a few hundred ms to deliver its first row, longer than the
        default timed region); rows are time-stamped on arrival and stop() keeps those between mark_begin and mark_end."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 - 0.05 <= t <= (self.t1 or time.time()) + 0.05]
        window = "timed region"
        if not inside:                       # timed region shorter than the sampling period: rows since the warm-up (same load)
            inside, window = [r for _, r in self.rows], "warm-up + timed region"
        for r in inside:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU baselines
def _ref_worker_init(max_n, max_l, out_prefix):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    ref = ref_loader.load_reference(max_n, max_l, out_prefix)
    S, NP = load_tables()
    ref.cfg.args.sub_scores, ref.cfg.args.np_scores = S, NP
    global _REF
    _REF = ref


def _ref_worker(read):
    _REF.bam.realign_read(read)
    return 1


def cpu_reference_run(reads, cores, out_prefix="/tmp/npore_bench_ref"):
    """The unmodified reference (oracle/_ref) driven like realign.py:110-114: Pool(cores).imap_unordered(realign_read)."""
    import multiprocessing as mp
    if os.path.exists(out_prefix + ".sam"):
        os.remove(out_prefix + ".sam")
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores, initializer=_ref_worker_init, initargs=(6, 100, out_prefix)) as pool:
        list(pool.imap_unordered(_ref_worker, reads[:cores], chunksize=1))          # warm the workers (import, tables)
        if os.path.exists(out_prefix + ".sam"):
            os.remove(out_prefix + ".sam")
        t0 = time.perf_counter()
        for _ in pool.imap_unordered(_ref_worker, reads, chunksize=max(1, min(100, len(reads) // (4 * cores) or 1))):
            pass
        dt = time.perf_counter() - t0
    return dt


def cpu_port_run(reads):
    """Fallback when oracle/_ref is absent: the single-threaded C restatement (oracle/npore_oracle.c)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    S, NP = load_tables()
    t0 = time.perf_counter()
    for r in reads:
        oracle.realign_cigar(r[9], r[7], r[5], S, NP)
    return time.perf_counter() - t0


def n_cu_of(reads, r=30, max_b_rows=20000):
    tot = 0
    for rd in reads:
        ops = len(rd[9]) + len(rd[7])
        nch = -(-ops // (max_b_rows - 1)) if ops else 0
        tot += (ops + nch) * (2 * r + 1)
    return tot


def file_e2e(ref, reads, S, NP, n=1000):
    """Informational (not the contract's e2e): the first n reads of the workload as a BAM FILE in, realigned SAM FILE out
    through npore_b200.bamio.realign_bam (native BGZF/BAM decode, GPU, native SAM text); never fails the bench line."""
    try:
        import re
        import tempfile
        from npore_b200 import bamio, cfg
        cfg.args.sub_scores, cfg.args.np_scores = S, NP
        with tempfile.TemporaryDirectory() as d:
            recs = [{"name": r[0], "flag": r[1], "ref_id": 0, "pos": r[3], "mapq": r[4], "seq": r[7], "qual": bytes([30] * len(r[7])),
                     "cigar": [(int(a), b) for a, b in re.findall(r"(\d+)(\D)", r[5])], "tags": {"HP": r[10]}}
                    for r in sorted(reads[:n], key=lambda r: r[3])]
            bamio.write_bam(os.path.join(d, "in.bam"), "@HD\tVN:1.6\tSO:coordinate\n", [("chr1", len(ref))], recs)
            fa = {"chr1": ref}
            bamio.realign_bam(os.path.join(d, "in.bam"), fa, out_prefix=os.path.join(d, "warm"), argv=["bench.py"], max_reads=32)
            best, phases = None, {}
            for _ in range(3):
                tm = {}
                t0 = time.perf_counter()
                got = bamio.realign_bam(os.path.join(d, "in.bam"), fa, out_prefix=os.path.join(d, "out"), argv=["bench.py"], timings=tm)
                dt = time.perf_counter() - t0
                if best is None or dt < best:
                    best, phases = dt, tm
            return {"reads": got, "reads_per_s": got / best, "seconds": best, "phase_seconds": {k: round(v, 4) for k, v in phases.items()},
                    "sam_bytes": os.path.getsize(os.path.join(d, "out.sam")), "host_threads": os.cpu_count()}
    except Exception as e:          # noqa: BLE001
        return {"unavailable": repr(e)}


def cpu_baseline(reads, budget_reads_per_core=16):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    cores = os.cpu_count() or 1
    try:
        import ref_loader
        have_ref = ref_loader.available()
    except Exception:
        have_ref = False
    if have_ref:
        sample = reads[:min(len(reads), max(64, budget_reads_per_core * cores))]
        dt = cpu_reference_run(sample, cores)
        kind, used = "reference", cores
    else:
        sample = reads[:24]
        dt = cpu_port_run(sample)
        kind, used = "port", 1
    cu = n_cu_of(sample)
    out = {"value": cu / dt / 1e9, "unit": "GCUPS", "cores": used, "kind": kind, "reads_per_s": len(sample) / dt,
           "sample": f"first {len(sample)} reads of the workload ({cu/1e9:.3f} GCU), wall {dt:.2f} s, "
                     + ("mp.Pool(imap_unordered, realign_read) over all host cores" if kind == "reference" else "1 thread, C port")}
    if have_ref:         # also the 1-process figure (a Pool of one worker, 6 reads)
        try:
            one = sample[:6]
            d1 = cpu_reference_run(one, 1)
            out["one_process"] = {"value": n_cu_of(one) / d1 / 1e9, "unit": "GCUPS", "reads_per_s": len(one) / d1}
        except Exception as e:      # noqa: BLE001
            out["one_process"] = {"unavailable": repr(e)}
    return out


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="npore_b200", choices=["npore_b200", "reference"])
    ap.add_argument("--reads", type=int, default=3000)
    ap.add_argument("--read-len", type=int, default=10000)
    ap.add_argument("--ref-len", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-file-e2e", action="store_true", help="skip the informational BAM-file-to-SAM-file measurement")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    S, NP = load_tables()
    config = {"workload": "C2: synthetic 1 Mb n-polymer-rich reference, 30x ONT-like 10 kb reads, per GPU "
                          f"({args.reads} reads x {args.read_len} bp over {args.ref_len} bp; seed 20260101+rank)",
              "params": "r=30 max_b_rows=20000 max_n=6 max_l=100 indel_start=5 indel_extend=1; align+standardise+collapse",
              "cache": "inputs larger than L2: 7.7 GB of traceback rows streamed per step, no reuse between steps",
              "parallelism": f"region shards x{world}, no collective"}

    if args.impl == "reference":
        # the reference's own CPU path on this box's host cores; rank 0 only
        if rank != 0:
            return
        _, reads = make_workload(20260101, args.ref_len, args.reads, args.read_len, NP)
        cores = os.cpu_count() or 1
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_loader
        kind = "reference" if ref_loader.available() else "port"
        per_step = max(32, min(len(reads), 8 * cores)) if kind == "reference" else 8
        times = []
        for s in range(args.warmup + args.steps):
            sample = reads[(s * per_step) % max(1, len(reads) - per_step):][:per_step]
            dt = cpu_reference_run(sample, cores) if kind == "reference" else cpu_port_run(sample)
            if s >= args.warmup:
                times.append((dt, n_cu_of(sample), len(sample)))
        tot_t = sum(t for t, _, _ in times); tot_cu = sum(c for _, c, _ in times); tot_r = sum(r for _, _, r in times)
        val = tot_cu / tot_t / 1e9
        line = {"impl": "reference", "metric": "GCUPS", "value": val, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "reads_per_s": tot_r / tot_t,
                "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": cores if kind == "reference" else 1, "kind": kind,
                                 "sample": f"{per_step} reads per step of the same workload"},
                "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: npore_b200 has no CPU fallback"}))
        sys.exit(2)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from npore_b200.engine import NPORE_OUT_NO_EXPANDED, NPORE_OUT_RLE, NPORE_OUT_STANDARDIZE, Realigner
    ref, reads = make_workload(20260101 + rank, args.ref_len, args.reads, args.read_len, NP)
    packed = pack_reads(reads, pinned=True)
    eng = Realigner(S, NP, device=local)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    flags = NPORE_OUT_STANDARDIZE | NPORE_OUT_RLE | NPORE_OUT_NO_EXPANDED
    result = eng.new_result(packed, flags, pinned=True)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- kernels only, inputs resident in HBM
    eng.upload(packed)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        eng.run(flags)
    sync_all()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fwd_ms, launches, stats = [], 0, None
    e0.record(stream)
    for _ in range(args.steps):
        eng.run(flags)
        stats = eng.stats()
        fwd_ms.append(stats["ms_forward"]); launches += stats["launches"]
    e1.record(stream)
    sync_all()
    dev_ms = e0.elapsed_time(e1)
    # ---------------- end to end through the C-ABI call with pinned host buffers
    for _ in range(max(1, args.warmup // 2)):
        eng.align_packed(packed, flags, result)
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    d2h = 0
    for _ in range(args.steps):
        eng.align_packed(packed, flags, result)
        d2h = eng.stats()["d2h_bytes"]
    e3.record(stream)
    sync_all()
    e2e_ms = e2.elapsed_time(e3)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None

    n_cu = stats["n_cu"]
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    u = torch.tensor([float(n_cu), float(len(reads))], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms = t.tolist()
    tot_cu, tot_reads = u.tolist()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = tot_cu * args.steps / (dev_ms * 1e-3) / 1e9
    e2e_val = tot_cu * args.steps / (e2e_ms * 1e-3) / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    fwd = float(np.mean(fwd_ms)) * 1e-3
    fwd_gcups = n_cu / fwd / 1e9
    achieved = n_cu * BYTES_PER_CU / fwd / 1e9
    alu_peak = stats["sm_count"] * 128 * sm_mhz * 1e6 / OPS_PER_CU / 1e9
    line = {
        "metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "reads_per_s": tot_reads * args.steps / (dev_ms * 1e-3),
        "e2e": {"value": e2e_val, "unit": "GCUPS", "reads_per_s": tot_reads * args.steps / (e2e_ms * 1e-3),
                "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": int(packed.h2d_bytes()), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "kernel_ms": {k: stats[k] for k in ("ms_plan", "ms_annotate", "ms_forward", "ms_traceback", "ms_finish", "ms_kernels_total")},
        "roofline": {"kernel": "forward_kernel<2>", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH if (args.reads, args.read_len) == (3000, 10000) else None,
                     "traffic_source": "profiles/r01_final_forward_metrics.txt (dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture of this workload)",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(n_cu * BYTES_PER_CU), "launch_ms": fwd * 1e3},
        "roofline_alu": {"bound": "cuda-core issue (SURVEY 8(d): 28 lane-ops per cell update)", "achieved": fwd_gcups,
                         "peak": alu_peak, "unit": "GCUPS", "frac": fwd_gcups / alu_peak},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(reads)
    elif world == 1:
        line["cpu_baseline"] = None
    if world == 1 and not args.no_file_e2e:
        line["bam_to_sam"] = file_e2e(ref, reads, S, NP)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
