#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: achieved HBM GB/s of sdot / saxpy / snrm2.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (oracle port) on host cores

One "step" = one pass of the hot path over one batch of synthetic fp32 data: sdot, saxpy, snrm2 at
n = 2^28 per GPU, unit stride (BASELINE.json's metric: "sdot/saxpy/snrm2 achieved HBM GB/s at n=2^28").
Algorithmic bytes per element: sdot 8 + saxpy 12 + snrm2 4 = 24.  Weak scaling: every
rank holds a 2^28-element block shard of each vector; saxpy needs no communication, each reduction ends
with one 48-byte record exchange.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_ELEM = {"sdot": 8, "sdsdot": 8, "sasum": 4, "snrm2": 4, "isamax": 4, "saxpy": 12, "sscal": 8, "scopy": 8,
                  "sswap": 16, "srot": 16, "srotm": 16}
STEP_ROUTINES = ("sdot", "saxpy", "snrm2")
STEP_BYTES_PER_ELEM = sum(BYTES_PER_ELEM[r] for r in STEP_ROUTINES)   # 24
NOMINAL_HBM_GBS = 8000.0          # the "~8 TB/s" of BASELINE.json's metric
FALLBACK_HBM_GBS = 6650.0         # /opt/skills/guides/B200_PROFILING.md, if MEASURED_PEAKS.json is absent
SAXPY_A = 1.000123


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_key: str):
    """dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU is under load (rank 0's GPU)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.1:
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except Exception:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# the reference arm: the oracle (C port of the Haskell code) on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_step_factory(n: int):
    """One CPU step = sdot + saxpy (pure: clone of y + update, Util.hs:134-142) + snrm2 over n elements."""
    import numpy as np
    from oracle import oracle
    oracle.build()
    rng = np.random.default_rng(1)
    x = rng.random(n, dtype=np.float32) * np.float32(2) - np.float32(1)
    y = rng.random(n, dtype=np.float32) * np.float32(2) - np.float32(1)
    out = np.empty_like(y)
    L, p = oracle.lib(), oracle._p

    def step():
        L.lbo_sdot(n, p(x), 1, p(y), 1)
        L.lbo_saxpy(n, SAXPY_A, p(x), 1, p(y), n, 1, p(out))
        L.lbo_snrm2(n, p(x), 1)
    return step


def time_cpu(n: int, steps: int, warmup: int):
    step = cpu_step_factory(n)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return dt, STEP_BYTES_PER_ELEM * n / dt / 1e9


def cpu_info():
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except Exception:
        pass
    return model, os.cpu_count()


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                      # rank 0 alone runs the CPU arm
    n = 1 << args.ref_log2n
    dt, gbs = time_cpu(n, args.steps, args.warmup)
    model, ncpu = cpu_info()
    sample = (f"each step = sdot+saxpy(pure: clone+update)+snrm2 on a 2^{args.ref_log2n}-element sample of the 2^{args.log2n} workload, "
              f"U(-1,1) fp32, unit stride; C port of Single.hs/Util.hs (gcc -O2, no FMA), 1 thread: the reference is single-threaded "
              f"pure Haskell (no ghc/stack in this image, so `stack bench` itself cannot run); host = {model}, {ncpu} logical CPUs")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 3), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_OUT = None


def emit(line: dict) -> None:
    out = _OUT if _OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "sdot+saxpy+snrm2 achieved HBM GB/s at n=2^28 per GPU (algorithmic bytes / device time)"


def workload_config(args, world: int) -> dict:
    n = 1 << args.log2n
    return {
        "workload": (f"sdot, saxpy, snrm2 fp32, unit stride, n=2^{args.log2n} per GPU, in place on device-resident block shards "
                     f"(BASELINE.json metric; configs[1]-[2] class); all ten routines are reported under `routines`"),
        "n_per_gpu": n, "n_global": n * world, "bytes_per_elem_step": STEP_BYTES_PER_ELEM,
        "l2": "each vector is 1 GiB per GPU, far larger than the 126 MB L2, so no flush between iterations",
        "parallelism": f"{world} rank(s), one per GPU; saxpy: no communication; sdot/snrm2: one 48-byte record exchange per call",
    }


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import numpy as np
    import torch
    import lambda_blas_b200 as lb
    from lambda_blas_b200 import inplace as ip, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = lb.Context(local)
    if world > 1:
        sharded.attach_torch_distributed(ctx)
        if args.exchange != "auto":
            ctx.set_fused_exchange(args.exchange == "fused")
    if args.pdl != 1:
        lb._ffi.check(lb._ffi.lib().lbk_set_pdl(ctx.handle, args.pdl), ctx.handle)
    for spec in args.launch:
        name, v, t, b = spec.split(":")
        ctx.set_launch(name, int(v), int(t), int(b))
    exchange = ("fused: record pushed into every peer's mailbox over NVLink inside the reduction kernel" if ctx.fused_exchange()
                else "ncclAllGather of one 48-byte record + finish kernel") if world > 1 else "none (1 rank)"

    n = 1 << args.log2n
    ng = n * world
    K, W = args.steps, args.warmup
    x, y = ctx.alloc_sharded(ng).fill_hash(1), ctx.alloc_sharded(ng).fill_hash(2)
    assert len(x) == n
    res = ctx.alloc(8)
    flags = lb.srotmg(1.5, 0.75, 2.0, 0.5)
    cs, sn = math.cos(0.3), math.sin(0.3)

    calls = {
        "sdot": lambda: ip.sdot_async(ng, x, 1, y, 1, res),
        "saxpy": lambda: ip.saxpy_(ng, SAXPY_A, x, 1, y, 1),
        "snrm2": lambda: ip.snrm2_async(ng, x, 1, res),
        "sdsdot": lambda: ip.sdsdot_async(ng, 0.5, x, 1, y, 1, res),
        "sasum": lambda: ip.sasum_async(ng, x, 1, res),
        "isamax": lambda: ip.isamax_async(ng, x, 1, res),
        "sscal": lambda: ip.sscal_(ng, SAXPY_A, x, 1),
        "scopy": lambda: ip.scopy_(ng, x, 1, y, 1),
        "sswap": lambda: ip.sswap_(ng, x, 1, y, 1),
        "srot": lambda: ip.srot_(ng, x, 1, y, 1, cs, sn),
        "srotm": lambda: ip.srotm_(flags, ng, x, 1, y, 1),
    }

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(values):
        if dist is None:
            return list(values)
        t = torch.tensor(list(values), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    sampler = ClockSampler(local) if rank == 0 else None
    t_load0 = time.time()

    # ---- the step: W warm-up, then exactly K timed steps bracketed by barrier + synchronize ----
    for _ in range(max(W, 3)):
        for r in STEP_ROUTINES:
            calls[r]()
    e0, e1 = ctx.event(), ctx.event()
    barrier()
    launches0, pdl0 = ctx.launch_count(), ctx.pdl_count()
    e0.record()
    for s in range(K):
        for r in STEP_ROUTINES:
            calls[r]()
    e1.record()
    barrier()
    launches, pdl_launches = ctx.launch_count() - launches0, ctx.pdl_count() - pdl0
    ms_total = e0.elapsed_ms(e1)
    step_check = res.to_host()   # reads the last snrm2 result: the step's output reaches the host

    # ---- the same K steps again with a CUDA event around every launch: per-kernel launch durations in
    # step context (the roofline's `achieved`).  Not part of `value`: an event between two kernels
    # forbids the programmatic overlap of a reduction's tail with the next kernel. ----
    marks = [[ctx.event() for _ in range(len(STEP_ROUTINES) + 1)] for _ in range(K)]
    barrier()
    for s in range(K):
        marks[s][0].record()
        for j, r in enumerate(STEP_ROUTINES):
            calls[r]()
            marks[s][j + 1].record()
    barrier()
    per_routine_in_step = {r: sum(marks[s][j].elapsed_ms(marks[s][j + 1]) for s in range(K)) / K
                           for j, r in enumerate(STEP_ROUTINES)}
    ms_step, saxpy_ms_in_step = max_over_ranks([ms_total / K, per_routine_in_step["saxpy"]])[:2]
    value = STEP_BYTES_PER_ELEM * ng / ms_step / 1e6

    # ---- every routine alone (K launches back to back), for the per-routine table ----
    routines = {}
    order = ["sdot", "saxpy", "snrm2", "sasum", "isamax", "sscal", "scopy", "sswap", "srot", "srotm", "sdsdot"]
    times = []
    for r in order:
        x.fill_hash(1); y.fill_hash(2)
        for _ in range(3):
            calls[r]()
        barrier()
        a, b = ctx.event().record(), None
        for _ in range(K):
            calls[r]()
        b = ctx.event().record()
        barrier()
        times.append(a.elapsed_ms(b) / K)
    times = max_over_ranks(times)
    peak, peak_src = measured_peak()
    for r, ms in zip(order, times):
        gbs = BYTES_PER_ELEM[r] * ng / ms / 1e6
        routines[r] = {"ms": round(ms, 4), "GBps": round(gbs, 1), "GBps_per_gpu": round(gbs / world, 1),
                       "pct_of_8TBs_per_gpu": round(100 * gbs / world / NOMINAL_HBM_GBS, 1),
                       "frac_of_measured_peak": round(gbs / world / peak, 3)}

    # ---- the Double instances (ddot, daxpy, ...: SURVEY.md 8f.3) on the same bytes: n/2 doubles per GPU ----
    routines_f64 = {}
    if not args.no_f64:
        nd, ngd = n // 2, (n // 2) * world
        xd, yd = ctx.alloc_sharded(ngd, np.float64).fill_hash(1), ctx.alloc_sharded(ngd, np.float64).fill_hash(2)
        resd = ctx.alloc(8, np.float64)
        dcalls = {
            "ddot": lambda: ip.ddot_async(ngd, xd, 1, yd, 1, resd), "daxpy": lambda: ip.daxpy_(ngd, SAXPY_A, xd, 1, yd, 1),
            "dnrm2": lambda: ip.dnrm2_async(ngd, xd, 1, resd), "dasum": lambda: ip.dasum_async(ngd, xd, 1, resd),
            "idamax": lambda: ip.idamax_async(ngd, xd, 1, resd), "dscal": lambda: ip.dscal_(ngd, SAXPY_A, xd, 1),
            "dcopy": lambda: ip.dcopy_(ngd, xd, 1, yd, 1), "dswap": lambda: ip.dswap_(ngd, xd, 1, yd, 1),
            "drot": lambda: ip.drot_(ngd, xd, 1, yd, 1, cs, sn), "drotm": lambda: ip.drotm_(flags, ngd, xd, 1, yd, 1),
        }
        dtimes = []
        for r, f in dcalls.items():
            xd.fill_hash(1); yd.fill_hash(2)
            for _ in range(3):
                f()
            barrier()
            a = ctx.event().record()
            for _ in range(K):
                f()
            b = ctx.event().record()
            barrier()
            dtimes.append(a.elapsed_ms(b) / K)
        dtimes = max_over_ranks(dtimes)
        for (r, _), ms in zip(dcalls.items(), dtimes):
            gbs = 2 * BYTES_PER_ELEM["s" + r[1:] if r != "idamax" else "isamax"] * ngd / ms / 1e6
            routines_f64[r] = {"ms": round(ms, 4), "GBps": round(gbs, 1), "pct_of_8TBs_per_gpu": round(100 * gbs / world / NOMINAL_HBM_GBS, 1)}
        del xd, yd, resd

    # ---- end to end: host buffers in, host results out, copies inside the timed region ----
    # The caller's x, y live in pinned HOST memory.  Per step: upload x and y in chunks; each chunk's saxpy goes
    # out of place into y' (lbk_saxpy_oop, so y stays intact for the reductions) as soon as the chunk has landed;
    # a second context's stream downloads y' chunk by chunk behind an event while the first keeps uploading
    # (H2D and D2H overlap on the two copy engines); then sdot and snrm2 over the whole (sharded) vectors, the
    # scalars come back, both streams are synchronised: the caller holds every result.
    x.fill_hash(1); y.fill_hash(2)
    hx, hy, hy_out, hres = lb.pinned_array(n), lb.pinned_array(n), lb.pinned_array(n), lb.pinned_array(8)
    x.download_async(hx); y.download_async(hy); ctx.sync()
    Ke = max(3, min(K, args.e2e_steps))
    ctx_dl = lb.Context(local)                      # same GPU, its own stream: the download lane
    yout = ctx.alloc(n)
    C = args.e2e_chunks
    m = n // C
    xs = [ctx.wrap(x.ptr + 4 * k * m, m, keepalive=x) for k in range(C)]
    ys = [ctx.wrap(y.ptr + 4 * k * m, m, keepalive=y) for k in range(C)]
    ws = [ctx.wrap(yout.ptr + 4 * k * m, m, keepalive=yout) for k in range(C)]
    wd = [ctx_dl.wrap(yout.ptr + 4 * k * m, m, keepalive=yout) for k in range(C)]
    evs = [ctx.event() for _ in range(C)]

    def e2e_step():
        for k in range(C):
            xs[k].upload_async(hx[k * m:(k + 1) * m]); ys[k].upload_async(hy[k * m:(k + 1) * m])
            ip.saxpy_into(m, SAXPY_A, xs[k], 1, ys[k], 1, ws[k])
            evs[k].record()
            ctx_dl.wait_event(evs[k])
            wd[k].download_async(hy_out[k * m:(k + 1) * m])
        ip.sdot_async(ng, x, 1, y, 1, res)
        ip.snrm2_async(ng, x, 1, res.view(2, 2))
        res.download_async(hres)
        ctx.sync(); ctx_dl.sync()                   # the caller holds the results here
    for _ in range(2):
        e2e_step()
    barrier()
    tw0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    barrier()
    e2e_ms = max_over_ranks([(time.perf_counter() - tw0) / Ke * 1e3])[0]
    e2e = {"value": round(STEP_BYTES_PER_ELEM * ng / e2e_ms / 1e6, 2), "unit": "GB/s", "ms_per_step": round(e2e_ms, 3), "steps": Ke,
           "h2d_bytes_per_step": 8 * n * world, "d2h_bytes_per_step": (4 * n + 32) * world, "chunks": C,
           "timing": "host wall clock around Ke steps, each ending in a synchronise of both streams; max over ranks",
           "api": ("pinned host x,y -> per chunk: lbk_vec_upload_async x2, lbk_saxpy_oop, event -> second context: lbk_wait_event, "
                   "lbk_vec_download_async(y') -> lbk_sdot_async, lbk_snrm2_async, download scalars -> lbk_sync x2")}
    ctx_dl.sync()
    t_load1 = time.time()
    clocks = sampler.stop(t_load0, t_load1) if sampler else None

    # ---- parity spot check of the timed configuration (size-independent property) ----
    # after the e2e step: y' = y + a*x exactly (separately rounded), checked on a slice against numpy fp32
    sl = slice(0, 1 << 16)
    want = (hy[sl] + np.float32(SAXPY_A) * hx[sl]).astype(np.float32)
    assert np.array_equal(hy_out[sl].view(np.uint32), want.view(np.uint32)), "saxpy result differs from the fp32 definition"

    if rank == 0:
        saxpy_gbs_gpu = BYTES_PER_ELEM["saxpy"] * n / saxpy_ms_in_step / 1e6
        roofline = {"bound": "hbm", "kernel": f"lbk::{ncu_traffic('saxpy_kernel') or 'map_unit_kernel<AxpyF<float>, 8, 2, 6, 0>'} (saxpy, the largest share of the step)",
                    "achieved": round(saxpy_gbs_gpu, 1), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "frac": round(saxpy_gbs_gpu / peak, 4), "frac_of_nominal_8TBs": round(saxpy_gbs_gpu / NOMINAL_HBM_GBS, 4),
                    "algorithmic_bytes_per_launch": BYTES_PER_ELEM["saxpy"] * n, "launch_ms": round(saxpy_ms_in_step, 4),
                    "launch_ms_how": "CUDA events around each launch, on the library's stream, in an instrumented repeat of the K timed steps",
                    "traffic": ncu_traffic("saxpy")}
        cpu_baseline = None
        if world == 1 and not args.no_cpu:
            cpu_steps = 20           # about 10 s of CPU work on one core
            dt, gbs = time_cpu(1 << args.ref_log2n, cpu_steps, 1)
            model, ncpu = cpu_info()
            cpu_baseline = {"value": round(gbs, 3), "unit": "GB/s", "cores": 1, "kind": "port",
                            "sample": (f"{cpu_steps} steps of sdot+saxpy(pure)+snrm2 on a 2^{args.ref_log2n}-element sample (the oracle, C port of the "
                                       f"reference's Haskell, single thread as the reference is); host {model}, {ncpu} logical CPUs")}
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": dict(workload_config(args, world), exchange=exchange),
            "pct_of_8TBs_per_gpu": round(100 * value / world / NOMINAL_HBM_GBS, 1),
            "frac_of_measured_peak_per_gpu": round(value / world / peak, 3),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "pdl_launches": int(pdl_launches), "clocks": clocks, "routines": routines, "routines_f64": routines_f64,
            "step_routines_ms": {r: round(v, 4) for r, v in per_routine_in_step.items()},
            "last_snrm2": float(step_check[0]),
        }
        emit(line)
    del xs, ys, ws, wd, evs
    ctx_dl.close()
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()



# ---------------------------------------------------------------------------------------------------
# --criterion: the reference's own benchmark matrix (test/Bench.hs), re-stated
# ---------------------------------------------------------------------------------------------------
# test/Bench.hs:17-113 times every Level-1 routine on 32 767-element vectors of `genEveryday` floats for
# n in {1..9, 10..90, 100..900, 1000..9000, 10000, 20000, 30000} x inc in {1, 10, 100} with (n-1)*inc < 32767
# (`vectorbench`, Bench.hs:116-132), naming each case `level-1/<routine>/<impl>/<routine>(n,inc)`.  Here
#   impl = pure   the oracle (C port of the reference's Haskell, one core) -- a PORT: no GHC in this image
#   impl = b200   the reference's call surface over device-resident vectors (lambda_blas_b200.single): scalar
#                 routines return the scalar to the host, vector routines return a NEW device vector (the
#                 reference's pure result) and the call is timed to completion on the stream
# plus BASELINE.json configs[0], `sdot(1000000,1)`: mean and standard deviation over >= 10 runs.  These sizes are
# launch-latency territory for a GPU (a 30 000-element sdot moves 240 KB): the table documents where the CPU
# wins; it is not the bandwidth benchmark.
NMAX = 32767


def criterion_sizes(quick):
    ns = list(range(1, 10)) + list(range(10, 100, 10)) + list(range(100, 1000, 100)) + list(range(1000, 10000, 1000)) + [10000, 20000, 30000]
    return [n for n in ns if not quick or n in (1, 10, 100, 1000, 10000, 30000)]


def criterion_measure(f, min_time=2e-3, min_runs=5):
    """criterion in spirit: run until min_time has elapsed (at least min_runs), report mean and stddev per call."""
    f(); f()
    ts = []
    t_end = time.perf_counter() + min_time
    while len(ts) < min_runs or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t0)
        if len(ts) >= 2000:
            break
    return statistics.mean(ts), (statistics.stdev(ts) if len(ts) > 1 else 0.0), len(ts)


def run_criterion(args) -> None:
    """--criterion: the reference's benchmark matrix (test/Bench.hs) for the CPU port and the B200 path, as CSV on stdout."""
    import numpy as np
    import lambda_blas_b200 as lb
    from oracle import oracle
    out = _OUT if _OUT is not None else sys.stdout
    quick = args.quick
    oracle.build()
    rng = np.random.default_rng(20260101)
    lo, hi = 2.0 ** -24, 2.0 ** 24 - 1                       # genEveryday, Gen.hs:208-214
    hv = [rng.uniform(lo, hi, NMAX).astype(np.float32) for _ in range(2)]
    ctx = lb.Context(0)
    dv = [ctx.to_device(h) for h in hv]
    a = np.float32(rng.uniform(lo, hi))
    c, s = np.float32(math.cos(0.3)), np.float32(math.sin(0.3))
    flags_o, flags_d = oracle.srotmg(1.5, 0.75, 2.0, 0.5), lb.srotmg(1.5, 0.75, 2.0, 0.5)
    x, y, X, Y = hv[0], hv[1], dv[0], dv[1]

    def done(r):            # a vector result is complete when the stream has run it
        ctx.sync()
        return r
    routines = {
        "sdot": (lambda n, i: oracle.sdot(n, x, i, y, i), lambda n, i: lb.sdot(n, X, i, Y, i)),
        "sasum": (lambda n, i: oracle.sasum(n, x, i), lambda n, i: lb.sasum(n, X, i)),
        "snrm2": (lambda n, i: oracle.snrm2(n, x, i), lambda n, i: lb.snrm2(n, X, i)),
        "isamax": (lambda n, i: oracle.isamax(n, x, i), lambda n, i: lb.isamax(n, X, i)),
        "sscal": (lambda n, i: oracle.sscal(n, a, x, i), lambda n, i: done(lb.sscal(n, a, X, i))),
        "scopy": (lambda n, i: oracle.scopy(n, x, i, y, i), lambda n, i: done(lb.scopy(n, X, i, Y, i))),
        "sswap": (lambda n, i: oracle.sswap(n, x, i, y, i), lambda n, i: done(lb.sswap(n, X, i, Y, i))),
        "saxpy": (lambda n, i: oracle.saxpy(n, a, x, i, y, i), lambda n, i: done(lb.saxpy(n, a, X, i, Y, i))),
        "srot": (lambda n, i: oracle.srot(n, x, i, y, i, c, s), lambda n, i: done(lb.srot(n, X, i, Y, i, c, s))),
        "srotm": (lambda n, i: oracle.srotm(flags_o, n, x, i, y, i), lambda n, i: done(lb.srotm(flags_d, n, X, i, Y, i))),
    }
    out.write("name,n,inc,mean_us,stddev_us,runs\n")
    for r, (cpu, gpu) in routines.items():
        for inc in (1, 10, 100):
            for n in criterion_sizes(quick):
                if (n - 1) * inc >= NMAX:
                    continue
                for impl, f in (("pure", cpu), ("b200", gpu)):
                    m, sd, k = criterion_measure(lambda: f(n, inc))
                    out.write(f"\"level-1/{r}/{impl}/{r}({n},{inc})\",{n},{inc},{m * 1e6:.2f},{sd * 1e6:.2f},{k}" + "\n"); out.flush()
    # BASELINE.json configs[0]: sdot n = 1e6, unit stride
    n = 1_000_000
    hx, hy = rng.uniform(lo, hi, n).astype(np.float32), rng.uniform(lo, hi, n).astype(np.float32)
    dx, dy = ctx.to_device(hx), ctx.to_device(hy)
    for impl, f in (("pure", lambda: oracle.sdot(n, hx, 1, hy, 1)), ("b200", lambda: lb.sdot(n, dx, 1, dy, 1))):
        m, sd, k = criterion_measure(f, min_time=0.2, min_runs=10)
        out.write(f"\"level-1/sdot/{impl}/sdot({n},1)\",{n},1,{m * 1e6:.2f},{sd * 1e6:.2f},{k}" + "\n"); out.flush()
    ctx.close()



def main():
    # Libraries (NCCL's version banner, torchrun notices) may print to stdout; the contract is ONE JSON
    # line there, so everything else goes to stderr and the line is written to the saved descriptor.
    global _OUT
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=28, help="log2 of the per-GPU vector length")
    ap.add_argument("--ref-log2n", type=int, default=26, help="log2 of the CPU sample length")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-chunks", type=int, default=16, help="chunks per vector in the end-to-end pipeline")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-f64", action="store_true", help="skip the table of the Double instances")
    ap.add_argument("--launch", action="append", default=[], metavar="ROUTINE:VARIANT:THREADS:BLOCKS_PER_SM",
                    help="override a routine's launch shape (lbk_set_launch), for A/B runs; the default shapes are the product")
    ap.add_argument("--pdl", type=int, default=1, help="lbk_set_pdl value (1 = the product's default; 0 = every kernel launched the ordinary way, for A/B runs)")
    ap.add_argument("--criterion", action="store_true", help="print the reference's criterion matrix (test/Bench.hs) for the CPU port and the B200 path as CSV")
    ap.add_argument("--quick", action="store_true", help="--criterion: six sizes per routine instead of the full list")
    ap.add_argument("--exchange", default="auto", choices=["auto", "fused", "nccl"],
                    help="how sharded reductions end: in-kernel over NVLink peer memory, or ncclAllGather + finish kernel")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.criterion:
        run_criterion(args)
        return
    if args.impl == "reference":
        run_reference(args)
        return
    if world != args.gpus and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    run_ours(args)


if __name__ == "__main__":
    main()
