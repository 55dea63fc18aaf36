#!/usr/bin/env python
"""bench.py -- throughput of the realignment hot path (nPoRe align(): aln.pyx:379-787 + per-read glue) on B200.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                    (the reference's own CPU implementation, host cores)

A "step" = one pass of the hot path over the workload.  Workloads are BASELINE.json's:
  N = 1   configs[1] ("C2"): seeded 1 Mb n-polymer-rich reference, 3,000 ONT-like 10 kb reads (30x), align() defaults
          (r=30, max_b_rows=20000) + CIGAR standardisation + collapse, i.e. what realign_read does per read.
  N >= 2  configs[2] ("C3"): 64 Mb reference, 192,000 reads, built as 64 seeded 1 Mb tiles (tile t = the C2 generator with
          seed 20260101 + 7919*(t+1), coordinates offset by t Mb), REGION-SHARDED: rank g realigns the reads of the
          contiguous genomic region [64g/N, 64(g+1)/N) Mb (equal cell-update load, reported per rank), every rank
          independently (no data-path collective), and the per-region CIGARs are gathered on the host in region order.
          Total work is fixed: STRONG scaling.  (--scaling strong --gpus 1 runs C3 on one GPU.)

Printed JSON line (rank 0): metric GCUPS (cell updates / s).
  value       kernels only, inputs resident in HBM, CUDA events on the launching stream (N=1: one bracket around K steps; C3: the
              library's per-batch CUDA-event time summed over the rank's batches, max over ranks)
  e2e         N=1: BAM FILE in -> realigned SAM FILE out through the package's public API npore_b200.bamio.realign_bam
              (native BGZF/BAM decode, H2D, kernels, D2H, native SAM text, file write all inside the timed region);
              C3: pinned host buffers -> H2D -> kernels -> D2H into ONE host buffer shared by the ranks, in region order
              (shard -> realign -> gather), device-timed from barrier to barrier, max over ranks
  e2e_packed  (N=1) the C-ABI call npore_align_batch on pre-packed pinned host buffers (H2D + kernels + D2H)
  roofline    the binding roof of the dominant kernel: CUDA-core instruction issue, 28 lane-ops per cell update;
              the HBM view (2 B per cell update against the measured copy bandwidth) is in roofline_hbm
A parity gate runs before anything is timed: the GPU's standardised CIGARs of a sample of the timed reads must equal the
CPU oracle's (n_parity_checked); a mismatch aborts the bench.
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
# forward_kernel<2> on the C2 batch, one ncu --set full capture (profiles/r02_forward_metrics.txt): dram read + write bytes per launch
NCU_DRAM_BYTES_PER_LAUNCH = 9.494e9
C3_TILES, TILE_REF, TILE_READS, READ_LEN = 64, 1_000_000, 3000, 10000
REF_CHUNKSIZE_NOTE = "imap_unordered chunksize = min(100, reads / (4 * cores)) (realign.py:110-114 uses 100; smaller here so that a bounded sample still spreads over all cores)"


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


def tile_seed(t):
    return 20260101 + 7919 * (t + 1)


def make_tile(t):
    """One 1 Mb tile of the C3 data set as flat arrays: ONE shared reference slice + per-read windows.
    Deterministic in t, so every rank can build exactly its own region of the same global read set."""
    from npore_b200.cig import bases_to_int
    from npore_b200.engine import bases_to_int_batch, cigars_to_rle_batch
    _, NP = load_tables()
    ref, reads = make_workload(tile_seed(t), TILE_REF, TILE_READS, READ_LEN, NP)
    seq_codes, seq_len = bases_to_int_batch([r[7] for r in reads])
    words, off = cigars_to_rle_batch([r[5] for r in reads])
    return {"tile": t, "ref_codes": bases_to_int(ref), "ref_start": np.array([r[3] for r in reads], np.int64),
            "ref_len": np.array([r[6] - r[3] for r in reads], np.int32), "seq_codes": seq_codes, "seq_len": seq_len,
            "cigar_rle": words, "cigar_off": off, "first_reads": reads[:8]}


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


# ------------------------------------------------------------------------------------------------ CPU side: baselines + parity oracle
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


def _oracle_cigar(read):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    S, NP = load_tables()
    return oracle.realign_cigar(read[9], read[7], read[5], S, NP)


def parity_gate(reads, gpu_cigars):
    """GPU standardised + collapsed CIGARs of `reads` against the CPU oracle (the checker; all host cores).  Returns the number
    of reads checked; raises SystemExit on the first mismatch -- nothing is timed on a wrong kernel."""
    import multiprocessing as mp
    cores = min(len(reads), os.cpu_count() or 1)
    with mp.get_context("spawn").Pool(cores) as pool:
        want = pool.map(_oracle_cigar, reads, chunksize=max(1, len(reads) // (4 * cores)))
    for k, (w, g) in enumerate(zip(want, gpu_cigars)):
        if w != g:
            print(json.dumps({"error": f"parity gate: GPU CIGAR of read {k} ({reads[k][0]}) differs from the CPU oracle", "oracle": w[:200], "gpu": g[:200]}))
            sys.exit(3)
    return len(reads)


def n_cu_of(reads, r=30, max_b_rows=20000):
    tot = 0
    for rd in reads:
        ops = len(rd[9]) + len(rd[7])
        nch = -(-ops // (max_b_rows - 1)) if ops else 0
        tot += (ops + nch) * (2 * r + 1)
    return tot


def write_fixture_bam(path, ref, reads):
    import re
    from npore_b200 import bamio
    recs = [{"name": r[0], "flag": r[1], "ref_id": 0, "pos": r[3], "mapq": r[4], "seq": r[7], "qual": bytes([30] * len(r[7])),
             "cigar": [(int(a), b) for a, b in re.findall(r"(\d+)(\D)", r[5])], "tags": {"HP": r[10]}}
            for r in sorted(reads, key=lambda r: r[3])]
    bamio.write_bam(path, "@HD\tVN:1.6\tSO:coordinate\n", [("chr1", len(ref))], recs)


def file_e2e(ref, reads, S, NP, steps, warmup, devices=None):
    """The public API end to end: a BAM FILE of the workload in, the realigned SAM FILE out (bamio.realign_bam: native BGZF/BAM
    decode, GPU, native SAM text, file write).  Returns seconds per step (wall, best is NOT taken: mean of the timed steps),
    host seconds per phase and the SAM size."""
    import tempfile
    from npore_b200 import bamio, cfg
    cfg.args.sub_scores, cfg.args.np_scores = S, NP
    with tempfile.TemporaryDirectory() as d:
        write_fixture_bam(os.path.join(d, "in.bam"), ref, reads)
        fa = {"chr1": ref}
        out = os.path.join(d, "out")
        times, phases, got = [], {}, 0
        for s in range(warmup + steps):
            tm = {}
            t0 = time.perf_counter()
            got = bamio.realign_bam(os.path.join(d, "in.bam"), fa, out_prefix=out, argv=["bench.py"], timings=tm, devices=devices)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
                for k, v in tm.items():
                    if isinstance(v, float):
                        phases[k] = phases.get(k, 0.0) + v / steps
        return {"reads": got, "seconds_per_step": float(np.mean(times)), "phase_seconds": {k: round(v, 4) for k, v in phases.items()},
                "bam_bytes": os.path.getsize(os.path.join(d, "in.bam")), "sam_bytes": os.path.getsize(out + ".sam"), "host_threads": os.cpu_count()}


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
                     + ("mp.Pool(imap_unordered, realign_read) over all host cores; " + REF_CHUNKSIZE_NOTE if kind == "reference" else "1 thread, C port")}
    if have_ref:         # also the 1-process figure (a Pool of one worker, 6 reads)
        try:
            one = sample[:6]
            d1 = cpu_reference_run(one, 1)
            out["one_process"] = {"value": n_cu_of(one) / d1 / 1e9, "unit": "GCUPS", "reads_per_s": len(one) / d1}
        except Exception as e:      # noqa: BLE001
            out["one_process"] = {"unavailable": repr(e)}
    return out


def peaks():
    p = {}
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    return (float(p.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in p else "fallback (B200_PROFILING.md)",
            float(p.get("sm_max_mhz", 1965.0)))


def roofline_blocks(n_cu, fwd_s, sm_count, c2_shape):
    hbm_peak, peak_src, sm_mhz = peaks()
    fwd_gcups = n_cu / fwd_s / 1e9
    alu_peak = sm_count * 128 * sm_mhz * 1e6 / OPS_PER_CU / 1e9
    achieved = n_cu * BYTES_PER_CU / fwd_s / 1e9
    traffic = NCU_DRAM_BYTES_PER_LAUNCH if c2_shape else None
    roof = {"kernel": "forward_kernel<2>", "bound": "cuda-core issue", "achieved": fwd_gcups, "peak": alu_peak, "unit": "GCUPS",
            "frac": fwd_gcups / alu_peak,
            "peak_source": f"{sm_count} SMs x 128 lanes x sm_max_mhz {sm_mhz:.0f} MHz ({peak_src.split(' ')[0]}) / {OPS_PER_CU} lane-ops per cell update (SURVEY 8(d))",
            "traffic": traffic, "traffic_unit": "bytes of DRAM read + write per launch (ncu --set full, profiles/r02_forward_metrics.txt)",
            "traffic_over_algorithmic": (traffic / (n_cu * BYTES_PER_CU)) if traffic else None,
            "algorithmic_bytes_per_launch": int(n_cu * BYTES_PER_CU), "launch_ms": fwd_s * 1e3}
    hbm = {"bound": "hbm (does not bind: 2 B per cell update)", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
           "peak_source": peak_src}
    return roof, hbm


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference_arm(args, config):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, NP = load_tables()
    seed = 20260101 if args.workload == "C2" else tile_seed(0)
    _, reads = make_workload(seed, TILE_REF, TILE_READS, READ_LEN, NP)
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
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "reads_per_s": tot_r / tot_t,
            "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": cores if kind == "reference" else 1, "kind": kind,
                             "sample": f"{per_step} reads per step of the same workload (the CPU path does not use the GPUs: the same host cores "
                                       f"at every --gpus N); " + REF_CHUNKSIZE_NOTE},
            "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ N = 1: C2
def run_c2(args, config, S, NP, local):
    import torch
    from npore_b200.engine import NPORE_OUT_NO_EXPANDED, NPORE_OUT_RLE, NPORE_OUT_STANDARDIZE, Realigner
    from npore_b200.cig import bases_to_int
    from npore_b200.engine import PackedBatch, bases_to_int_batch, cigars_to_rle_batch
    ref, reads = make_workload(20260101, args.ref_len, args.reads, args.read_len, NP)
    # packed host buffers as a caller holding a coordinate-sorted region has them: ONE shared reference slice + per-read windows
    seq_codes, seq_len = bases_to_int_batch([r[7] for r in reads])
    words, off = cigars_to_rle_batch([r[5] for r in reads])
    packed = PackedBatch.from_flat_shared(bases_to_int(ref), np.array([r[3] for r in reads], np.int64), np.array([r[6] - r[3] for r in reads], np.int32),
                                          seq_codes, seq_len, words, off)
    for name in ("ref_codes", "seq_codes", "cigar_rle"):          # pinned staging
        tt = torch.from_numpy(np.array(getattr(packed, name))).pin_memory()
        setattr(packed, "_pin_" + name, tt)
        setattr(packed, name, tt.numpy())
    eng = Realigner(S, NP, device=local)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    flags = NPORE_OUT_STANDARDIZE | NPORE_OUT_RLE | NPORE_OUT_NO_EXPANDED
    result = eng.new_result(packed, flags, pinned=True)

    # ---------------- parity gate (before anything is timed)
    n_check = 0 if args.no_parity_gate else min(len(reads), args.parity_reads)
    eng.align_packed(packed, flags, result)
    if n_check:
        n_check = parity_gate(reads[:n_check], [result.cigar_text(i) for i in range(n_check)])

    # ---------------- kernels only, inputs resident in HBM
    eng.upload(packed)
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        eng.run(flags)
    torch.cuda.synchronize()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fwd_ms, launches, stats = [], 0, None
    e0.record(stream)
    for _ in range(args.steps):
        eng.run(flags)
        stats = eng.stats()
        fwd_ms.append(stats["ms_forward"]); launches += stats["launches"]
    e1.record(stream)
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1)
    # ---------------- the C-ABI call with pre-packed pinned host buffers
    for _ in range(max(1, args.warmup // 2)):
        eng.align_packed(packed, flags, result)
    torch.cuda.synchronize()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    d2h = 0
    for _ in range(args.steps):
        eng.align_packed(packed, flags, result)
        d2h = eng.stats()["d2h_bytes"]
        launches += eng.stats()["launches"]
    e3.record(stream)
    torch.cuda.synchronize()
    packed_ms = e2.elapsed_time(e3)
    eng.close()
    # ---------------- the public API: BAM file -> SAM file
    fe = None
    if not args.no_file_e2e:
        try:
            fe = file_e2e(ref, reads, S, NP, args.steps, max(1, args.warmup // 2))
            launches += 0      # (realign_bam's own launches are not added to the count of the two device-timed regions above)
        except Exception as e:      # noqa: BLE001
            fe = {"unavailable": repr(e)}
    sampler.mark_end()
    clocks = sampler.stop()

    n_cu = stats["n_cu"]
    n_reads = len(reads)
    value = n_cu * args.steps / (dev_ms * 1e-3) / 1e9
    packed_val = n_cu * args.steps / (packed_ms * 1e-3) / 1e9
    fwd = float(np.mean(fwd_ms)) * 1e-3
    roof, hbm = roofline_blocks(n_cu, fwd, stats["sm_count"], (args.reads, args.read_len) == (3000, 10000))
    e2e_packed = {"value": packed_val, "unit": "GCUPS", "reads_per_s": n_reads * args.steps / (packed_ms * 1e-3), "ms_per_step": packed_ms / args.steps,
                  "h2d_bytes_per_step": int(packed.h2d_bytes()), "d2h_bytes_per_step": int(d2h),
                  "what": "npore_align_batch on pre-packed pinned host buffers (one shared reference slice, 1 B/base read codes, run-length CIGARs): "
                          "H2D + kernels + D2H (CUDA events)"}
    if fe and "seconds_per_step" in fe:
        e2e = {"value": n_cu / fe["seconds_per_step"] / 1e9, "unit": "GCUPS", "reads_per_s": fe["reads"] / fe["seconds_per_step"],
               "ms_per_step": 1e3 * fe["seconds_per_step"],
               "h2d_bytes_per_step": int(len(ref) + sum(len(r[7]) for r in reads) + 4 * int(packed.cigar_off[-1])), "d2h_bytes_per_step": int(d2h),
               "what": "public API bamio.realign_bam: BAM file in (native BGZF inflate + record decode), one shared reference slice + reads H2D, kernels, "
                       "run-length CIGARs D2H, native SAM text, SAM file out; wall clock around the call",
               "phase_seconds": fe["phase_seconds"], "bam_bytes": fe["bam_bytes"], "sam_bytes": fe["sam_bytes"], "host_threads": fe["host_threads"]}
    else:
        e2e = dict(e2e_packed, note="file pipeline unavailable: " + str(fe))
    line = {
        "metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "reads_per_s": n_reads * args.steps / (dev_ms * 1e-3),
        "e2e": e2e, "e2e_packed": e2e_packed,
        "gpu_launches": int(launches), "n_parity_checked": int(n_check),
        "kernel_ms": {k: stats[k] for k in ("ms_plan", "ms_annotate", "ms_forward", "ms_traceback", "ms_finish", "ms_kernels_total")},
        "roofline": roof, "roofline_hbm": hbm, "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(reads)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ C3: region-sharded, strong scaling
def run_c3(args, config, S, NP, rank, world, local):
    import multiprocessing as mp
    import torch
    import torch.distributed as dist
    from npore_b200.engine import (NPORE_OUT_NO_EXPANDED, NPORE_OUT_RLE, NPORE_OUT_STANDARDIZE, BatchResult, PackedBatch, PipelinedRealigner,
                                   Realigner, rle_to_text)
    flags = NPORE_OUT_STANDARDIZE | NPORE_OUT_RLE | NPORE_OUT_NO_EXPANDED
    n_tiles = args.tiles
    mine = list(range(n_tiles * rank // world, n_tiles * (rank + 1) // world))           # contiguous genomic region of this rank
    t0 = time.perf_counter()
    procs = max(1, min(len(mine), (os.cpu_count() or 1) // world))
    with mp.get_context("spawn").Pool(procs) as pool:
        tiles = pool.map(make_tile, mine, chunksize=1)
    synth_s = time.perf_counter() - t0
    packs = [PackedBatch.from_flat_shared(t["ref_codes"], t["ref_start"], t["ref_len"], t["seq_codes"], t["seq_len"], t["cigar_rle"], t["cigar_off"])
             for t in tiles]
    for p in packs:     # pinned staging, as a caller streaming a BAM would hold it
        for name in ("ref_codes", "seq_codes", "cigar_rle"):
            a = getattr(p, name)
            tt = torch.from_numpy(np.array(a)).pin_memory()
            setattr(p, "_pin_" + name, tt)
            setattr(p, name, tt.numpy())
    h2d = sum(p.h2d_bytes() for p in packs)

    # ---- the gather target: ONE host buffer shared by all ranks (POSIX shared memory), tile slots in region order
    cap_words = args.tile_rle_cap
    shm_path = f"/dev/shm/npore_bench_{os.environ.get('MASTER_PORT', '0')}_{world}"
    nbytes = n_tiles * (cap_words + TILE_READS + 1 + 15) * 4 * 2
    if rank == 0:
        with open(shm_path, "wb") as fh:
            fh.truncate(nbytes)
    if world > 1:
        dist.barrier()
    shm = torch.from_file(shm_path, shared=True, size=nbytes // 4, dtype=torch.int32)
    try:
        torch.cuda.cudart().cudaHostRegister(shm.data_ptr(), nbytes, 0)
    except Exception:       # noqa: BLE001
        pass
    shm_np = shm.numpy()
    slot_words = nbytes // 4 // n_tiles

    def slot(t):
        base = t * slot_words
        rle = shm_np[base:base + cap_words].view(np.uint32)
        off = shm_np[base + cap_words:base + cap_words + 2 * (TILE_READS + 1)].view(np.int64)
        return rle, off

    eng = Realigner(S, NP, device=local)
    n_cu = 0
    # ---------------- parity gate on the first reads of the rank's first tile
    n_check = 0
    res0 = eng.align_packed(packs[0], flags, eng.new_result(packs[0], flags, pinned=False))
    if not args.no_parity_gate and rank == 0:
        fr = tiles[0]["first_reads"]
        n_check = parity_gate(fr, [rle_to_text(res0.rle_words(i)) for i in range(len(fr))])
    # ---------------- kernels only: every tile uploaded once, then W untimed + K timed runs (CUDA events inside the library)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    kern_ms = fwd_ms = 0.0
    launches = 0
    kernel_ms = {}
    for p in packs:
        eng.upload(p)
        for _ in range(args.warmup):
            eng.run(flags)
        if p is packs[0]:
            sampler.mark_begin()
        for _ in range(args.steps):
            eng.run(flags)
            st = eng.stats()
            kern_ms += st["ms_kernels_total"]; fwd_ms += st["ms_forward"]; launches += st["launches"]
            for k in ("ms_plan", "ms_annotate", "ms_forward", "ms_traceback", "ms_finish", "ms_kernels_total"):
                kernel_ms[k] = kernel_ms.get(k, 0.0) + st[k] / args.steps
        n_cu += st["n_cu"]
    sm_count = st["sm_count"]
    eng.close()
    # ---------------- shard -> realign -> gather: two batches in flight per GPU, results straight into the shared host buffer
    pipe = PipelinedRealigner(S, NP, n_inflight=2, device=local)
    results = []
    for t, p in zip(mine, packs):
        rle, off = slot(t)
        ops = p.ref_len.astype(np.int64) + p.seq_len
        r = BatchResult(p.n, 0, int((-(-ops // 19999)).sum()) + 8, True, pinned=False, want_ops=False, rle_buf=rle, rle_off_buf=off[:p.n + 1])
        results.append(r)

    def one_pass():
        futs = [pipe.submit(p, flags, result=r) for p, r in zip(packs, results)]
        d2h, nl = 0, 0
        for f in futs:
            _, st_ = f.result()
            d2h += st_["d2h_bytes"]; nl += st_["launches"]
        return d2h, nl

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(1, args.warmup // 2)):
        one_pass()
    sync_all()
    stream = torch.cuda.current_stream()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    own_s = 0.0
    d2h = 0
    gathered = 0
    for _ in range(args.steps):
        t1 = time.perf_counter()
        d2h, nl = one_pass()
        launches += nl
        own_s += time.perf_counter() - t1
        if world > 1:
            dist.barrier()                                   # every region's CIGARs are in the shared buffer
        if rank == 0:                                        # rank 0 walks the gathered buffer in region order
            gathered = sum(int(slot(t)[1][TILE_READS]) for t in range(n_tiles))
    e3.record(stream)
    sync_all()
    e2e_ms = e2.elapsed_time(e3)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    pipe.close()

    t = torch.tensor([kern_ms, e2e_ms, own_s * 1e3, fwd_ms], dtype=torch.float64, device="cuda")
    u = torch.tensor([float(n_cu), float(len(mine) * TILE_READS), float(h2d), float(d2h), float(launches)], dtype=torch.float64, device="cuda")
    lo = torch.tensor([float(n_cu), own_s * 1e3, synth_s], dtype=torch.float64, device="cuda")
    hi = lo.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    kern_ms, e2e_ms, own_ms, fwd_ms = t.tolist()
    tot_cu, tot_reads, h2d_all, d2h_all, launches_all = u.tolist()
    if rank == 0:
        os.remove(shm_path)
    if rank != 0:
        return
    value = tot_cu * args.steps / (kern_ms * 1e-3) / 1e9
    e2e_val = tot_cu * args.steps / (e2e_ms * 1e-3) / 1e9
    roof, hbm = roofline_blocks(tot_cu / world, fwd_ms / args.steps * 1e-3, sm_count, False)
    roof["note"] = "per GPU: the slowest rank's forward-kernel time for its region"
    line = {
        "metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": kern_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "reads_per_s": tot_reads * args.steps / (kern_ms * 1e-3),
        "e2e": {"value": e2e_val, "unit": "GCUPS", "reads_per_s": tot_reads * args.steps / (e2e_ms * 1e-3), "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                "what": "shard -> realign -> gather: per rank its region's batches (one shared reference slice + reads, pinned host) through "
                        "npore_align_batch, two in flight; run-length CIGARs D2H straight into ONE host buffer shared by the ranks, in region order; "
                        "barrier; rank 0 walks the gathered buffer.  Device-timed barrier to barrier, max over ranks",
                "gathered_rle_words": int(gathered),
                "slowest_rank_own_ms_per_step": own_ms / args.steps, "fastest_rank_own_ms_per_step": lo[1].item() / args.steps,
                "limiter": "the slowest rank's own pipeline (H2D + kernels + D2H); the gather itself is the D2H target, so it adds only the barrier wait "
                           f"= {(e2e_ms - own_ms) / args.steps:.2f} ms per step"},
        "shards": {"regions": world, "tiles_per_rank": n_tiles // world, "cell_updates_min": lo[0].item(), "cell_updates_max": hi[0].item(),
                   "synthesis_seconds_max": hi[2].item()},
        "gpu_launches": int(launches_all), "n_parity_checked": int(n_check),
        "kernel_ms": {k: round(v, 3) for k, v in kernel_ms.items()},
        "roofline": roof, "roofline_hbm": hbm, "clocks": clocks, "cpu_baseline": None,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="npore_b200", choices=["npore_b200", "reference"])
    ap.add_argument("--scaling", default="auto", choices=["auto", "strong"], help="auto: C2 on one GPU, region-sharded C3 on several; strong: C3 always")
    ap.add_argument("--reads", type=int, default=TILE_READS)
    ap.add_argument("--read-len", type=int, default=READ_LEN)
    ap.add_argument("--ref-len", type=int, default=TILE_REF)
    ap.add_argument("--tiles", type=int, default=C3_TILES, help="1 Mb tiles of the C3 data set (64 = 64 Mb, 192,000 reads)")
    ap.add_argument("--tile-rle-cap", type=int, default=1_600_000, help="run-length words reserved per tile in the gather buffer")
    ap.add_argument("--parity-reads", type=int, default=192)
    ap.add_argument("--no-parity-gate", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-file-e2e", action="store_true", help="skip the BAM-file-to-SAM-file e2e (e2e then repeats e2e_packed)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    S, NP = load_tables()
    c3 = world > 1 or args.gpus > 1 or args.scaling == "strong"
    args.workload = "C3" if c3 else "C2"
    params = "r=30 max_b_rows=20000 max_n=6 max_l=100 indel_start=5 indel_extend=1; align+standardise+collapse"
    if c3:
        config = {"workload": f"C3: synthetic chr20-scale reference ({args.tiles} Mb as {args.tiles} seeded 1 Mb n-polymer-rich tiles), 30x ONT-like 10 kb reads "
                              f"({args.tiles * TILE_READS} reads), region-sharded over {max(world, 1)} B200 by genomic coordinate (equal cell-update load)",
                  "params": params,
                  "cache": "inputs larger than L2: 7.7 GB of traceback rows streamed per 3,000-read batch, no reuse between batches or steps",
                  "parallelism": f"{max(world, 1)} contiguous genomic regions, one per GPU, no data-path collective; host gather in region order",
                  "note": "N=1 default runs BASELINE configs[1] (C2, one tile of the same generator); GCUPS is a rate, so the per-N values compare directly"}
    else:
        config = {"workload": "C2: synthetic 1 Mb n-polymer-rich reference, 30x ONT-like 10 kb reads, 1 B200 "
                              f"({args.reads} reads x {args.read_len} bp over {args.ref_len} bp; seed 20260101)",
                  "params": params,
                  "cache": "inputs larger than L2: 7.7 GB of traceback rows streamed per step, no reuse between steps",
                  "parallelism": "1 GPU (N >= 2 runs the region-sharded C3 set: strong scaling)"}

    if args.impl == "reference":
        run_reference_arm(args, config)
        return

    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: npore_b200 has no CPU fallback"}))
        sys.exit(2)
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if c3:
        run_c3(args, config, S, NP, rank, world, local)
    else:
        run_c2(args, config, S, NP, local)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
