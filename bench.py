#!/usr/bin/env python
"""bench.py -- throughput of the FaQCs trim + filter + statistics hot path on B200.

One "step" = one pass of the hot path (frame -> trim/filter/stats -> route/emit)
over one batch of synthetic Illumina-like reads (BASELINE.json configs[1]:
2x150 PE, ASCII-33, default BWA_plus trim + filters + full statistics).

  python bench.py --gpus N --steps K --warmup W          # this implementation
  python bench.py --impl reference ...                    # the reference's CPU path (oracle/_ref/FaQCs)

Prints ONE JSON line (rank 0).  `value` = reads/s with the batch resident in HBM
(CUDA events on the launching stream, max over ranks); `e2e` = the same metric
through fq_process_host with pinned HOST buffers (H2D + kernels + D2H inside the
timed region); `roofline` = algorithmic bytes of the dominant kernel / its
device time against the measured HBM copy peak; `cpu_baseline` = the reference
binary timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "reads/s trim+filter+stats (2x150 PE, default BWA_plus -q 5 --min_L 50 -n 2 --lc 0.85, full stats)"
UNIT = "reads/s"
WORKLOAD = "C2: synthetic Illumina 2x150 PE, ASCII-33, default trim + filters + full stats matrices"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "FaQCs")


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons WHILE the timed region runs (NVML, ~1 kHz; nvidia-smi as fallback)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.sm, self.sm_max, self.reasons = [], 0, set()
        self.source = "nvml"

    def _run_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        while not self.stop_flag.is_set():
            self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            try:
                r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for bit, name in self.REASONS.items():
                if r & bit:
                    self.reasons.add(name)
            self.stop_flag.wait(0.001)

    def _run_smi(self):
        self.source = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                f = [x.strip() for x in out.split(",")]
                self.sm.append(float(f[0]))
                self.sm_max = max(self.sm_max, float(f[1]))
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def result(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max or None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def time_reference_binary(w, threads: int, tmp_root: str):
    """Wall-clock of the unmodified reference on (r1, r2) with -t threads, --trim_only (skips only R)."""
    d = tempfile.mkdtemp(prefix="faqcs_bench_", dir=tmp_root)
    try:
        p1, p2 = os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq")
        w.r1.tofile(p1)
        w.r2.tofile(p2)
        t0 = time.perf_counter()
        p = subprocess.run([REF_BIN, "-1", p1, "-2", p2, "-d", os.path.join(d, "out"), "-t", str(threads), "--trim_only"],
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        dt = time.perf_counter() - t0
        if p.returncode != 0:
            raise RuntimeError("reference failed: " + p.stderr.decode(errors="replace")[-300:])
        return dt
    finally:
        shutil.rmtree(d, ignore_errors=True)


def tmp_root():
    return "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()


def cpu_baseline(sample_pairs: int):
    from faqcs_b200 import synth
    cores = os.cpu_count() or 1
    w = synth.c2(sample_pairs, start=7_000_000)
    if os.path.exists(REF_BIN):
        dt = time_reference_binary(w, cores, tmp_root())
        kind = "reference"
        sample = f"{sample_pairs} pairs (2x150) of the same workload, FaQCs v2.10 -t {cores} --trim_only, files on tmpfs, wall clock of the process"
    else:                                   # the reference binary did not travel: time the oracle port
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from faqcs_b200.api import Options
        from oracle_binding import OracleEngine
        with OracleEngine(Options(input_quality_offset=33)) as eng:
            t0 = time.perf_counter()
            eng.process(w.r1, w.r2)
            dt = time.perf_counter() - t0
        kind, cores = "port", 1
        sample = f"{sample_pairs} pairs (2x150), single-thread oracle port, in-memory"
    return {"value": 2 * sample_pairs / dt, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
            "gbases_per_s": 2 * sample_pairs * 150 / dt / 1e9}


def workload_config(args, world=1):
    """The `config` object both arms print: BASELINE configs[1] as this bench runs it."""
    reps = max(1, args.batch_pairs // args.block_pairs)
    batch_pairs = reps * args.block_pairs
    return {"workload": WORKLOAD, "pairs_per_step_per_gpu": batch_pairs * args.batches_per_step, "read_length": 150,
            "device_batch_pairs": batch_pairs, "batches_per_step": args.batches_per_step}


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (unmodified FaQCs v2.10 built from
    /root/reference into oracle/_ref) on all of this box's cores.  Same metric / unit / config as the b200 arm; each step is
    a bounded sample of that workload (args.ref_pairs pairs through the binary), so that W + K steps end within minutes."""
    if rank != 0:
        return
    from faqcs_b200 import synth
    cores = os.cpu_count() or 1
    pairs = args.ref_pairs
    w = synth.c2(pairs, start=9_000_000)
    d = tempfile.mkdtemp(prefix="faqcs_refarm_", dir=tmp_root())
    times = []
    try:
        p1, p2 = os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq")
        w.r1.tofile(p1)
        w.r2.tofile(p2)
        kind = "reference" if os.path.exists(REF_BIN) else "port"
        for i in range(args.warmup + args.steps):
            if kind == "reference":
                out = os.path.join(d, "out%d" % i)
                t0 = time.perf_counter()
                p = subprocess.run([REF_BIN, "-1", p1, "-2", p2, "-d", out, "-t", str(cores), "--trim_only"],
                                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
                dt = time.perf_counter() - t0
                if p.returncode != 0:
                    raise RuntimeError("reference failed: " + p.stderr.decode(errors="replace")[-300:])
                shutil.rmtree(out, ignore_errors=True)
            else:                               # the reference binary did not travel: time the oracle port
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                from faqcs_b200.api import Options
                from oracle_binding import OracleEngine
                with OracleEngine(Options(input_quality_offset=33)) as eng:
                    t0 = time.perf_counter()
                    eng.process(w.r1, w.r2)
                    dt = time.perf_counter() - t0
                cores = 1
            if i >= args.warmup:
                times.append(dt)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    total = sum(times)
    value = 2 * pairs * len(times) / total
    cfg = workload_config(args)
    cfg["gbases_per_s"] = value * 150 / 1e9
    sample = (f"each step = {pairs} pairs (2x150) of the same workload through FaQCs v2.10 -t {cores} --trim_only "
              f"(files on tmpfs, wall clock of the process); the reference streams at a size-independent rate")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


class _DevPtr:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def bind_to_gpu_numa_node(local_rank):
    """Several ranks per host: run this rank (and allocate its pinned buffers) on the CPUs next to its GPU, so the
    end-to-end leg's H2D / D2H traffic does not cross sockets.  Returns the previous affinity (restored before the
    reference binary is timed on all cores) or None when NVML / the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local_rank]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        old = os.sched_getaffinity(0)
        cpus &= old
        if cpus:
            os.sched_setaffinity(0, cpus)
            return old
    except Exception:
        pass
    return None


def stats_arrays(st):
    from faqcs_b200.api import Stats
    return {f: np.asarray(getattr(st, f)).astype(np.int64) for f in Stats.FIELDS}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--block-pairs", type=int, default=250_000, help="pairs generated on the host (numpy)")
    ap.add_argument("--batch-pairs", type=int, default=2_000_000, help="pairs per device batch (block replicated in HBM)")
    ap.add_argument("--batches-per-step", type=int, default=50,
                    help="device batches per step: 50 x 2 M pairs = the 100 M pairs of BASELINE configs[1] (the resident batch is replayed, SURVEY 8(d))")
    ap.add_argument("--contexts-per-gpu", type=int, default=2,
                    help="contexts (one batch in flight each, own stream, own host thread) sharing a GPU in the device-resident leg: "
                         "consecutive batches overlap on the device (one context's k_trim with another's framing / emit)")
    ap.add_argument("--e2e-steps", type=int, default=None, help="end-to-end steps (default: min(steps, 4); 0 disables)")
    ap.add_argument("--cpu-pairs", type=int, default=500_000, help="sample size of the cpu_baseline leg")
    ap.add_argument("--ref-pairs", type=int, default=500_000, help="pairs per step of --impl reference (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--read-length", type=int, default=150, help="c2 recipe at another read length (profiling only; the headline is 150)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2 is the headline config (BASELINE configs[1]); the others are the parity configs, for profiling only")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from faqcs_b200 import synth
    from faqcs_b200.api import Engine, Options

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: faqcs_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity0 = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes to stdout; the contract is ONE JSON line there.  NCCL_DEBUG=VERSION (this image's default) prints
        # the version banner straight to stdout, so that level is dropped; other levels are sent to stderr.
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            del os.environ["NCCL_DEBUG"]
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    # ---- workload: each rank owns its own slice of the read stream (weak scaling, no data-path collective)
    reps = max(1, args.batch_pairs // args.block_pairs)
    batch_pairs = reps * args.block_pairs
    nb = max(1, args.batches_per_step)
    from faqcs_b200.api import BUILTIN_ADAPTERS, MODE_HARD, POLYA_ADAPTER
    gen = {"c2": synth.c2, "c3": synth.c3, "c4": synth.c4, "c5": synth.c5}[args.workload]
    w = gen(args.block_pairs, start=rank * args.block_pairs, L=args.read_length) if args.workload == "c2" and args.read_length != 150 \
        else gen(args.block_pairs, start=rank * args.block_pairs)
    paired = w.r2 is not None
    d_r1 = torch.from_numpy(w.r1).to(dev).repeat(reps)
    d_r2 = torch.from_numpy(w.r2).to(dev).repeat(reps) if paired else torch.zeros(16, dtype=torch.uint8, device=dev)
    n1, n2 = d_r1.numel(), (d_r2.numel() if paired else 0)
    reads_per_batch = (2 if paired else 1) * batch_pairs
    reads_per_step = reads_per_batch * nb
    opts = {"c2": Options(),
            "c3": Options(filter_adapter=True, adapters=list(BUILTIN_ADAPTERS) + [POLYA_ADAPTER] + list(w.artifacts or [])),
            "c4": Options(qc_only=True),
            "c5": Options(mode=MODE_HARD, quality=20, average_quality=25.0, replace_to_N_q=10, discard_output=True)}[args.workload]
    n_ctx = max(1, args.contexts_per_gpu)
    engines = [Engine(opts, device=local_rank) for _ in range(n_ctx)]
    for e in engines:
        e.autodetect(w.r1, w.r2)
    eng = engines[0]
    ext = torch.cuda.ExternalStream(eng.stream(), device=dev)
    exts = [torch.cuda.ExternalStream(e.stream(), device=dev) for e in engines]

    def batch(e=eng):
        return e.process_device(d_r1.data_ptr(), n1, d_r2.data_ptr() if paired else None, n2, 0, True, copy_out=False)

    def run_batches(total, seg=None):
        """`total` batches dealt out to the contexts in order, each context on its own host thread (the C ABI call blocks
        until its batch is done; ctypes releases the GIL); returns when all are done."""
        errs = []

        def work(j):
            try:
                for _ in range(j, total, n_ctx):
                    batch(engines[j])
                    if seg is not None:
                        t = engines[j].last_timing()
                        for k in seg[j]:
                            seg[j][k] += t[k]
            except BaseException as ex:          # surfaced below: a failed batch must fail the bench
                errs.append(ex)

        if n_ctx == 1:
            work(0)
        else:
            th = [threading.Thread(target=work, args=(j,)) for j in range(n_ctx)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        if errs:
            raise errs[0]

    # one batch alone: its statistics are the unit the final check multiplies
    res = batch()
    one = stats_arrays(eng.stats())
    n_batches_done = 1
    for e in engines[1:]:
        batch(e)
        n_batches_done += 1
    for _ in range(max(args.warmup, 3)):
        run_batches(nb)
        n_batches_done += nb
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys = int(vis.split(",")[local_rank]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local_rank
    sampler = ClockSampler(phys)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = sum(e.launch_count() for e in engines)
    seg_ctx = [{k: 0.0 for k in ("all", "frame", "adapter", "trim", "emit")} for _ in engines]
    torch.cuda.synchronize()
    e0.record(ext)                                   # the device is idle: every context's stream starts after this point
    run_batches(args.steps * nb, seg_ctx)
    for x in exts[1:]:
        ext.wait_stream(x)                           # the end event follows the last kernel of every context
    e1.record(ext)
    torch.cuda.synchronize()
    n_batches_done += args.steps * nb
    if world > 1:
        dist.barrier()
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    launches = sum(e.launch_count() for e in engines) - launches0
    seg_overlapped = {k: sum(s_[k] for s_ in seg_ctx) for k in seg_ctx[0]}
    # per-kernel durations: with several contexts the events of one context's segment also span the other contexts' kernels,
    # so the kernels are timed once more on ONE context alone (same batch, after the timed region)
    iso_batches = min(nb, 20)
    seg = {k: 0.0 for k in seg_ctx[0]}
    if n_ctx > 1:
        for _ in range(iso_batches):
            batch()
            t = eng.last_timing()
            for k in seg:
                seg[k] += t[k]
        n_batches_done += iso_batches
    else:
        seg, iso_batches = seg_overlapped, args.steps * nb
    for e in engines[1:]:                            # contexts of one device: dst += src (fq_merge_stats)
        eng.merge_stats_from(e)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    value = world * reads_per_step * args.steps / (total_ms / 1e3)

    # ---- what the timed steps produced: every statistic must be exactly (batches run) x (one batch)
    st_local = eng.stats()
    got = stats_arrays(st_local)
    for f, a in one.items():
        if a.shape != got[f].shape or not np.array_equal(a * n_batches_done, got[f]):
            raise SystemExit(f"bench: statistics after {n_batches_done} batches are not {n_batches_done} x one batch ({f})")
    checked = f"statistics after {n_batches_done} identical batches == {n_batches_done} x one batch (all {len(one)} arrays)"

    import ctypes as C
    from faqcs_b200.api import CBatchOut
    out_bytes = list(res.stream_bytes)
    alg_batch = n1 + n2 + sum(out_bytes)          # B_in + B_out of one device batch

    # ---- multi-GPU merge of the statistics: the path's only collective (ncclAllReduce over NVLink inside the library)
    allreduce_ms = None
    if world > 1:
        from faqcs_b200 import dist_stats
        allreduce_ms = dist_stats.allreduce_engine_stats(eng, dist, dev)
        merged = stats_arrays(eng.stats())
        fs = torch.tensor(got["filter_stats"], device=dev)
        dist.all_reduce(fs, op=dist.ReduceOp.SUM)
        if not np.array_equal(fs.cpu().numpy(), merged["filter_stats"]):
            raise SystemExit("bench: merged filter counters differ from the sum of the ranks' counters")
        checked += "; merged block == sum over ranks"
    st = eng.stats()

    # ---- end-to-end: pinned host buffers in, host buffers out
    e2e = None
    if args.e2e_steps is None:
        args.e2e_steps = max(1, min(args.steps, 4))
    if args.e2e_steps > 0 and args.workload == "c2":
        h1, h2 = eng.host_alloc(n1), eng.host_alloc(n2)
        for k in range(reps):
            h1[k * w.r1.size:(k + 1) * w.r1.size] = w.r1
            h2[k * w.r2.size:(k + 1) * w.r2.size] = w.r2
        cb2 = CBatchOut()
        # pieces mode (fq_set_output_pieces): the streams come back as pieces of THESE pinned input buffers plus the literal
        # bytes of the records that changed -- what a writer hands to writev(2); checked against byte mode below
        eng.set_output_pieces(True)

        def pipeline(n_b):
            """submit(i+1); run(i); wait(i-1): upload, kernels and download of three consecutive batches overlap."""
            tk = [None] * n_b
            t = C.c_uint64()
            eng._check(eng.lib.fq_submit_host(eng.ctx, C.c_void_p(h1.ctypes.data), n1, C.c_void_p(h2.ctypes.data), n2, 0, 1, C.byref(t)))
            tk[0] = t.value
            for i in range(n_b):
                if i + 1 < n_b:
                    eng._check(eng.lib.fq_submit_host(eng.ctx, C.c_void_p(h1.ctypes.data), n1, C.c_void_p(h2.ctypes.data), n2, 0, 1, C.byref(t)))
                    tk[i + 1] = t.value
                eng._check(eng.lib.fq_run(eng.ctx, tk[i]))
                if i > 0:
                    eng._check(eng.lib.fq_wait(eng.ctx, tk[i - 1], C.byref(cb2)))
            eng._check(eng.lib.fq_wait(eng.ctx, tk[n_b - 1], C.byref(cb2)))

        pipeline(3)                                  # warm-up: allocates the pinned output slots
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(ext)
        pipeline(args.e2e_steps * nb)
        f1.record(ext)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ems = torch.tensor([max(f0.elapsed_time(f1), wall * 1e3)], device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        d2h = int(sum(int(cb2.literal_bytes[i]) + 16 * int(cb2.n_pieces[i]) for i in range(4)))
        # the pieces of the last timed batch must expand to exactly the byte-mode streams of the same batch
        import hashlib
        last = eng._collect(cb2, True)
        expanded = last.expand(h1, h2)
        eng.set_output_pieces(False)
        ref = eng.process(h1, h2)
        if [hashlib.sha256(x).digest() for x in expanded] != [hashlib.sha256(x).digest() for x in ref.streams]:
            raise SystemExit("bench: pieces of the end-to-end leg do not expand to the byte-mode streams")
        e2e = {"value": world * reads_per_step * args.e2e_steps / (float(ems.item()) / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": (n1 + n2) * nb, "d2h_bytes_per_step": d2h * nb, "steps": args.e2e_steps,
               "ms_per_step": float(ems.item()) / args.e2e_steps,
               "output_bytes_per_step": int(sum(int(cb2.bytes[i]) for i in range(4))) * nb,
               "api": "fq_submit_host / fq_run / fq_wait in pieces mode (pinned host buffers in; out: each stream as pieces of those "
                      "buffers + literal bytes of the changed records, expansion verified against byte mode after the timed region; "
                      "H2D, kernels and D2H of consecutive batches overlap)"}
        eng.host_free(h1)
        eng.host_free(h2)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        steps_batches = args.steps * nb
        ms_batch = total_ms / steps_batches
        # fraction = sum over reads of (B_in + B_out) / device time / peak -- the WHOLE path
        achieved = alg_batch / (ms_batch / 1e3) / 1e9
        # per-segment figures, each against ITS OWN algorithmic bytes: framing reads B_in; the trim/filter/stats kernel reads
        # the sequence and quality lines (2L per read) -- or, when emission is fused into it, B_in's share + B_out --;
        # a separate emit kernel reads B_in and writes B_out
        total_len = float(one["filter_stats"][2])
        seg_ms = {k: v / iso_batches for k, v in seg.items()}
        fused = seg_ms["emit"] <= 0.0 and sum(out_bytes) > 0
        seg_bytes = {"frame": n1 + n2, "trim": (n1 + n2 + sum(out_bytes)) if fused else 2 * total_len + 8 * reads_per_batch,
                     "emit": 0 if fused else n1 + n2 + sum(out_bytes), "adapter": total_len}
        traffic = {}
        try:        # measured DRAM bytes per input byte of each kernel (ncu capture of this code, profiles/r2_traffic.json)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            traffic = {k: v["dram_bytes_per_input_byte"] * (n1 + n2) for k, v in tj.items() if isinstance(v, dict)}
        except Exception:
            pass
        kernels = {}
        for k in ("frame", "adapter", "trim", "emit"):
            if seg_ms[k] > 0.01:            # an idle segment (no adapter pass, emission fused or off) is two events apart
                a = seg_bytes[k] / (seg_ms[k] / 1e3) / 1e9
                kernels[k] = {"ms": seg_ms[k], "ms_in_timed_region": seg_overlapped[k] / steps_batches,
                              "algorithmic_bytes": seg_bytes[k], "achieved": a, "frac": a / peak,
                              "dram_bytes": traffic.get(k), "traffic_ratio": (traffic[k] / seg_bytes[k]) if k in traffic and seg_bytes[k] else None}
        dom = max(kernels, key=lambda k: kernels[k]["ms"]) if kernels else None
        cfg = workload_config(args, world)
        rl = float(one["filter_stats"][2]) / max(float(one["filter_stats"][1]), 1.0)
        if args.workload != "c2":
            cfg["workload"] = args.workload + " (parity config, not the headline)"
        elif args.read_length != 150:
            cfg["workload"] = "c2 recipe at read length %d (profiling only, not the headline)" % args.read_length
        cfg.update({"read_length": rl, "bytes_in_per_batch": n1 + n2, "bytes_out_per_batch": sum(out_bytes),
                    "gbases_per_s": rl * value / 1e9, "l2": "inputs (%.0f MB per device batch) exceed the 126 MB L2" % ((n1 + n2) / 1e6),
                    "reads_total": int(st.filter_stats[1]), "reads_kept": int(st.filter_stats[3]),
                    "stats_allreduce_ms": allreduce_ms, "result_check": checked,
                    "contexts_per_gpu": n_ctx,
                    "kernel_timing": ("kernels[].ms: CUDA events of one context running alone (%d batches after the timed region); "
                                      "ms_in_timed_region: the same events while %d contexts share the device" % (iso_batches, n_ctx))
                                     if n_ctx > 1 else "kernels[].ms: CUDA events inside the timed region"})
        try:        # un-profiled bench lines of the other BASELINE configs, measured with this code (profiles/)
            cfg["other_workloads"] = json.load(open(os.path.join(ROOT, "profiles", "r2_other_workloads.json")))
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": cfg,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": sum(traffic.get(k, 0) for k in kernels) or None,
                         "kernel": "whole path (frame -> trim/filter/stats -> emit), per device batch", "ms_per_batch": ms_batch,
                         "algorithmic_bytes": alg_batch, "peak_source": peak_src, "dominant_kernel": dom, "kernels": kernels},
            "clocks": sampler.result(),
            "gpu_launches": int(launches),
        }
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:      # the CPU baseline is a 1-GPU-run item (rank 0, N = 1 only)
            if affinity0:
                os.sched_setaffinity(0, affinity0)
            line["cpu_baseline"] = cpu_baseline(args.cpu_pairs)
        print(json.dumps(line))
    for e in engines:
        e.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
