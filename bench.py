#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: achieved HBM GB/s of sdot / saxpy / snrm2.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (oracle port) on host cores

One "step" = PASSES (default 12) passes of the hot path over one batch of synthetic fp32 data, a pass being
sdot, saxpy, snrm2 at n = 2^28 per GPU, unit stride (BASELINE.json's metric: "sdot/saxpy/snrm2 achieved HBM
GB/s at n=2^28"); twelve passes make a step ~10.7 ms, so the K = 20 timed steps cover > 0.2 s of device time.
Algorithmic bytes per element and pass: sdot 8 + saxpy 12 + snrm2 4 = 24.  `value` is WEAK
scaling: every rank holds a 2^28-element block shard of each vector; saxpy needs no communication, each
reduction ends with one 48-byte record exchange.  With more than one rank the line also carries
  strong         the metric's n as the GLOBAL n: n_global = 2^28 and 2^30 (BASELINE.json configs[2]),
                 sdot/saxpy/snrm2/sasum/isamax, with the speed-up over ONE GPU of the same box on the same n;
  parity_check   an untimed block comparing the sharded path with the oracle / exact properties on this very job;
  single_process the same step driven from ONE host thread through lbk_init_multi (rank 0 alone, all GPUs).
Rank 0 prints ONE JSON line.
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


def time_cpu_all_cores(log2n_per_thread: int = 24, steps: int = 2):
    """The same CPU step on EVERY logical CPU at once, one independent copy per thread on its own vectors: what a caller of the
    (single-threaded, pure) reference could get by splitting the work by hand.  Reported beside the one-core figure, never as
    `value`: the reference itself uses one core."""
    import threading
    T = os.cpu_count() or 1
    n = 1 << log2n_per_thread
    steps_fns = [cpu_step_factory(n) for _ in range(T)]
    for f in steps_fns[:1]:
        f()
    start = threading.Barrier(T + 1)
    done = threading.Barrier(T + 1)

    def work(f):
        f()                                 # warm: touch this thread's pages
        start.wait()
        for _ in range(steps):
            f()                             # ctypes releases the GIL for the duration of each call
        done.wait()
    ths = [threading.Thread(target=work, args=(f,)) for f in steps_fns]
    for t in ths:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    done.wait()
    dt = time.perf_counter() - t0
    for t in ths:
        t.join()
    return {"value": round(STEP_BYTES_PER_ELEM * n * T * steps / dt / 1e9, 2), "unit": "GB/s", "threads": T,
            "sample": f"one copy of the step per logical CPU, 2^{log2n_per_thread} elements each, {steps} steps",
            "note": "hand-split over all cores; the reference is single-threaded, so `value` stays the one-core figure"}


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


def reference_toolchain() -> dict:
    """if the reference's toolchain ever shows up on the box, say so (and `__graft_entry__.build()`
    then builds oracle/_ref from /root/reference).  On this image every entry is None."""
    import shutil
    return {t: shutil.which(t) for t in ("stack", "ghc", "cabal", "gfortran")}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                      # rank 0 alone runs the CPU arm
    n = 1 << args.ref_log2n
    dt, gbs = time_cpu(n, args.steps, args.warmup)
    model, ncpu = cpu_info()
    tools = reference_toolchain()
    sample = (f"each step = sdot+saxpy(pure: clone+update)+snrm2 on a 2^{args.ref_log2n}-element sample of the 2^{args.log2n} workload "
              f"(the metric is a rate, so the sample carries), U(-1,1) fp32, unit stride; C port of Single.hs/Util.hs (gcc -O2, no FMA), "
              f"1 thread: the reference is single-threaded pure Haskell; toolchain probe {tools} -> `stack bench` cannot run here; "
              f"host = {model}, {ncpu} logical CPUs")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 3), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": 1, "kind": "port", "sample": sample,
                         "all_cores": time_cpu_all_cores()},
        "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "reference_toolchain": tools,
    }
    emit(line)


_OUT = None


def emit(line: dict) -> None:
    out = _OUT if _OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "sdot+saxpy+snrm2 achieved HBM GB/s at n=2^28 per GPU (algorithmic bytes / device time)"


def workload_config(args, world: int) -> dict:
    """Identical in both arms (it depends on the command line only), so the driver can tell they ran the same workload."""
    n = 1 << args.log2n
    return {
        "workload": (f"sdot, saxpy, snrm2 fp32, unit stride, n=2^{args.log2n} per GPU, in place on device-resident block shards "
                     f"(BASELINE.json metric; configs[1]-[2] class); all ten routines are reported under `routines`"),
        "n_per_gpu": n, "n_global": n * world, "bytes_per_elem_pass": STEP_BYTES_PER_ELEM,
        "passes_per_step": args.passes,
        "l2": f"each vector is {4 * n >> 20} MiB per GPU, far larger than the 126 MB L2, so no flush between iterations",
        "parallelism": f"{world} rank(s), one per GPU; saxpy: no communication; sdot/snrm2: one 48-byte record exchange per call",
        "cpu_sample": (f"the reference arm (and cpu_baseline) time the same three routines on a 2^{args.ref_log2n}-element sample per step, "
                       f"one pass per step, on one host core"),
    }


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def time_calls(ctx, barrier, f, k: int, warm: int = 3) -> float:
    """ms per call of f(i) over k back-to-back calls on the library's stream (CUDA events), after `warm` untimed ones."""
    for i in range(warm):
        f(i)
    barrier()
    a = ctx.event().record()
    for i in range(k):
        f(i)
    b = ctx.event().record()
    barrier()
    return a.elapsed_ms(b) / k


def planted(lb, np, ctx, n_global: int, world: int, rank: int, seed: int, plant: dict):
    """A sharded hash-filled vector with `plant` = {global index: value} written over it (each rank writes what falls into its block)."""
    from lambda_blas_b200 import sharded
    v = ctx.alloc_sharded(n_global).fill_hash(seed)
    lo, hi = sharded.shard_range(n_global, world, rank)
    for k, val in plant.items():
        if lo <= k < hi:
            ctx.wrap(v.ptr + 4 * (k - lo), 1, keepalive=v).upload(np.array([val], dtype=np.float32))
    return v


def parity_block(lb, np, ctx, world: int, rank: int, agree, log2_big: int) -> dict:
    """Untimed: the sharded path of THIS job against the oracle (moderate n) and against exact / double-precision properties at
    the full size of BASELINE.json configs[2].  `agree(values)` -> True when every rank passed the same numbers (all-ranks-same-bits,
    test/Main.hs:37-42 is the property the tests replay; here across GPUs).  Returns {"status": "ok" | "FAILED: ...", ...}."""
    from lambda_blas_b200 import inplace as ip, sharded
    from oracle import oracle
    oracle.build()
    fails, checks = [], 0

    def check(name, cond):
        nonlocal checks
        checks += 1
        if not cond:
            fails.append(name)

    # ---- moderate n: the whole (unsharded) vector through the oracle on every rank
    n = (1 << 22) + 3
    lo, hi = sharded.shard_range(n, world, rank)
    edge = sharded.shard_range(n, world, 0)[1]                 # first element of shard 1 (== n when there is one rank)
    e1 = min(edge, n - 1)
    cases = {
        "max at the end of shard 0": {edge - 1: -9.0},
        "max at the start of shard 1": {e1: 9.0},
        "tie across the shard boundary": {edge - 1: 7.0, e1: -7.0, n - 1: 7.0},
        "NaN at index 0": {0: float("nan"), e1: 5.0},
        "NaN elsewhere": {e1: float("nan"), edge - 1: 3.0},
    }
    u = 2.0 ** -24
    hy = lb.hash_fill_reference(n, 12)
    y = ctx.alloc_sharded(n).fill_hash(12)
    for name, plant in cases.items():
        hx = lb.hash_fill_reference(n, 11)
        for k, v in plant.items():
            hx[k] = v
        x = planted(lb, np, ctx, n, world, rank, 11, plant)
        im = lb.isamax(n, x, 1)
        check(f"isamax [{name}]", im == oracle.isamax(n, hx, 1))
        vals = [float(im)]
        if not np.isnan(hx).any():
            prod = float(np.sum(np.abs(hx.astype(np.float64) * hy)))
            d, nr, sa = lb.sdot(n, x, 1, y, 1), lb.snrm2(n, x, 1), lb.sasum(n, x, 1)
            check(f"sdot [{name}]", abs(float(d) - float(oracle.sdot(n, hx, 1, hy, 1))) <= 2 * n * u * prod)
            check(f"snrm2 [{name}]", abs(float(nr) - float(oracle.snrm2(n, hx, 1))) <= 2 * n * u * float(oracle.snrm2(n, hx, 1)))
            check(f"sasum [{name}]", abs(float(sa) - float(oracle.sasum(n, hx, 1))) <= 2 * n * u * float(np.sum(np.abs(hx))))
            vals += [float(d), float(nr), float(sa)]
        check(f"all ranks hold the same bits [{name}]", agree(vals))
    hx = lb.hash_fill_reference(n, 11)
    x = ctx.alloc_sharded(n).fill_hash(11)
    ip.saxpy_(n, SAXPY_A, x, 1, y, 1)
    want = oracle.saxpy(n, SAXPY_A, hx, 1, hy, 1)
    check("saxpy bit-exact on this rank's slice", np.array_equal(y.to_host().view(np.uint32), want[lo:hi].view(np.uint32)))
    del x, y

    # ---- BASELINE.json configs[2] at full size (n_global = 2^log2_big): exact values, double-precision cross-checks, planted extrema
    N = 1 << log2_big
    lo, hi = sharded.shard_range(N, world, rank)
    ones = ctx.alloc_sharded(N).fill(1.0)
    half = ctx.alloc_sharded(N).fill(0.5)
    check("sasum(0.5 x 2^k) exact", float(lb.sasum(N, half, 1)) == N / 2)
    check("sdot(1, 0.5) exact", float(lb.sdot(N, ones, 1, half, 1)) == N / 2)
    check("snrm2(1 x 2^k) exact", float(lb.snrm2(N, ones, 1)) == float(np.float32(math.sqrt(N))))
    check("isamax(all equal) = 0", lb.isamax(N, ones, 1) == 0)
    del half
    x, yv = ctx.alloc_sharded(N).fill_hash(21), ctx.alloc_sharded(N).fill_hash(22)
    pos = ctx.alloc_sharded(N).fill_hash(23, 0.0, 1.0)
    wide = lambda a, b: float(lb.sdsdot(N, 0.0, a, 1, b, 1))          # noqa: E731  products and sum in double (Single.hs:231-262)
    d, nr, sa = float(lb.sdot(N, x, 1, yv, 1)), float(lb.snrm2(N, x, 1)), float(lb.sasum(N, pos, 1))
    nx2, ny2 = wide(x, x), wide(yv, yv)
    # tolerance: 2^-18 of ||x|| ||y|| (>= sum|x_i y_i|) -- far inside the stated k n eps sum|x_i y_i|, which is > 1 at this n
    check("sdot vs double-precision accumulation", abs(d - wide(x, yv)) <= 2.0 ** -18 * math.sqrt(nx2 * ny2))
    check("snrm2 vs sqrt(double-precision sum of squares)", abs(nr - math.sqrt(nx2)) <= 2.0 ** -20 * math.sqrt(nx2))
    check("sasum vs double-precision sum", abs(sa - wide(pos, ones)) <= 2.0 ** -18 * wide(pos, ones))
    check("all ranks hold the same bits [2^%d]" % log2_big, agree([d, nr, sa, nx2]))
    del pos, ones
    for name, plant, want in (("max at the last element of shard 0", {sharded.shard_range(N, world, 0)[1] - 1: 3.0}, sharded.shard_range(N, world, 0)[1] - 1),
                              ("tie: last element of the vector and first of the last shard", {N - 1: -4.0, sharded.shard_range(N, world, world - 1)[0]: 4.0},
                               sharded.shard_range(N, world, world - 1)[0]),
                              ("NaN at 0 wins", {0: float("nan"), N - 1: 5.0}, 0)):
        xp = planted(lb, np, ctx, N, world, rank, 21, plant)
        im = lb.isamax(N, xp, 1)
        check(f"isamax 2^{log2_big} [{name}]", im == want and agree([float(im)]))
        del xp
    # saxpy at full size: bit-exact windows at both ends of this rank's block against the fp32 definition
    ip.saxpy_(N, SAXPY_A, x, 1, yv, 1)
    w = min(1 << 16, hi - lo)
    for off in (0, (hi - lo) - w):
        got = np.empty(w, dtype=np.float32)
        ctx.wrap(yv.ptr + 4 * off, w, keepalive=yv).download_async(got); ctx.sync()
        wx, wy = lb.hash_fill_reference(w, 21, offset=lo + off), lb.hash_fill_reference(w, 22, offset=lo + off)
        want_w = (wy + (np.float32(SAXPY_A) * wx).astype(np.float32)).astype(np.float32)
        check(f"saxpy 2^{log2_big} window @{off} bit-exact", np.array_equal(got.view(np.uint32), want_w.view(np.uint32)))
    del x, yv
    ok = agree([float(len(fails))]) and not fails
    return {"status": "ok" if ok else "FAILED: " + "; ".join(fails[:6]), "checks": checks, "ranks": world,
            "covers": (f"n = 2^22+3 vs the oracle on the whole vector: isamax with maxima / ties at shard edges and both NaN rules, sdot / snrm2 / sasum within "
                       f"2 n u sum|.|, saxpy bit-exact per slice; n_global = 2^{log2_big} (BASELINE.json configs[2]): exact sums of constants, sdot / snrm2 / sasum "
                       f"against double-precision accumulation (sdsdot), planted extrema at shard edges, saxpy bit-exact windows; every result identical on all ranks")}


def run_ours(args) -> None:
    import numpy as np
    import torch
    import lambda_blas_b200 as lb
    from lambda_blas_b200 import inplace as ip, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist, idle = None, None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idle = dist.new_group(backend="gloo")        # a CPU-side barrier: ranks waiting there leave their GPU idle
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
    K, W, P = args.steps, max(args.warmup, 3), args.passes
    x, y = ctx.alloc_sharded(ng).fill_hash(1), ctx.alloc_sharded(ng).fill_hash(2)
    assert len(x) == n
    res = ctx.alloc(8)
    flags = lb.srotmg(1.5, 0.75, 2.0, 0.5)
    cs, sn = math.cos(0.3), math.sin(0.3)

    def make_calls(xv, yv, nn, out):
        return {
            "sdot": lambda: ip.sdot_async(nn, xv, 1, yv, 1, out),
            "saxpy": lambda: ip.saxpy_(nn, SAXPY_A, xv, 1, yv, 1),
            "snrm2": lambda: ip.snrm2_async(nn, xv, 1, out),
            "sdsdot": lambda: ip.sdsdot_async(nn, 0.5, xv, 1, yv, 1, out),
            "sasum": lambda: ip.sasum_async(nn, xv, 1, out),
            "isamax": lambda: ip.isamax_async(nn, xv, 1, out),
            "sscal": lambda: ip.sscal_(nn, SAXPY_A, xv, 1),
            "scopy": lambda: ip.scopy_(nn, xv, 1, yv, 1),
            "sswap": lambda: ip.sswap_(nn, xv, 1, yv, 1),
            "srot": lambda: ip.srot_(nn, xv, 1, yv, 1, cs, sn),
            "srotm": lambda: ip.srotm_(flags, nn, xv, 1, yv, 1),
        }
    calls = make_calls(x, y, ng, res)

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

    def agree(values) -> bool:
        """True when every rank passes bit-identical numbers (NaN == NaN)."""
        if dist is None:
            return True
        a = torch.tensor(np.asarray(values, dtype=np.float64).view(np.int64), device="cuda")
        lo_t, hi_t = a.clone(), a.clone()
        dist.all_reduce(lo_t, op=dist.ReduceOp.MIN); dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
        return bool(torch.equal(lo_t, hi_t))

    sampler = ClockSampler(local) if rank == 0 else None
    t_load0 = time.time()

    # ---- the step: W warm-up steps, then exactly K timed steps (of P passes each) bracketed by barrier + synchronize ----
    def one_step():
        for _ in range(P):
            for r in STEP_ROUTINES:
                calls[r]()
    for _ in range(W):
        one_step()
    e0, e1 = ctx.event(), ctx.event()
    barrier()
    launches0, pdl0 = ctx.launch_count(), ctx.pdl_count()
    e0.record()
    for s in range(K):
        one_step()
    e1.record()
    barrier()
    launches, pdl_launches = ctx.launch_count() - launches0, ctx.pdl_count() - pdl0
    ms_total = e0.elapsed_ms(e1)
    step_check = res.to_host()   # reads the last snrm2 result: the step's output reaches the host

    # ---- the same passes again with a CUDA event around every launch: per-kernel launch durations in
    # step context (the roofline's `achieved`).  Not part of `value`: an event between two kernels
    # forbids the programmatic overlap of a reduction's tail with the next kernel. ----
    KI = min(K * P, 60)
    marks = [[ctx.event() for _ in range(len(STEP_ROUTINES) + 1)] for _ in range(KI)]
    barrier()
    for s in range(KI):
        marks[s][0].record()
        for j, r in enumerate(STEP_ROUTINES):
            calls[r]()
            marks[s][j + 1].record()
    barrier()
    per_routine_in_step = {r: sum(marks[s][j].elapsed_ms(marks[s][j + 1]) for s in range(KI)) / KI
                           for j, r in enumerate(STEP_ROUTINES)}
    del marks
    ms_step, saxpy_ms_in_step = max_over_ranks([ms_total / K, per_routine_in_step["saxpy"]])[:2]
    value = STEP_BYTES_PER_ELEM * ng * P / ms_step / 1e6

    # ---- every routine alone (KR launches back to back), for the per-routine table ----
    # These tables are about the HBM roofline of each kernel, so the elementwise kernels walk FORWARDS here (lbk_set_traversal 0):
    # by default they start at the end of the vectors the previous kernel left in the 126 MB L2, and back-to-back launches of one
    # routine on the same two vectors would then find ~3 % of their reads cached (`routines_l2_walk` below has those figures).
    # The timed step above ran with the library's default.
    ctx.set_traversal(0)
    KR = max(K, 20)
    peak, peak_src = measured_peak()

    def table(call_map, order, nn_global, bytes_of):
        out = {}
        times = []
        for r in order:
            x.fill_hash(1); y.fill_hash(2)
            times.append(time_calls(ctx, barrier, lambda i, f=call_map[r]: f(), KR))
        for r, ms in zip(order, max_over_ranks(times)):
            gbs = bytes_of(r) * nn_global / ms / 1e6
            out[r] = {"ms": round(ms, 4), "GBps": round(gbs, 1), "GBps_per_gpu": round(gbs / world, 1),
                      "pct_of_8TBs_per_gpu": round(100 * gbs / world / NOMINAL_HBM_GBS, 1),
                      "frac_of_measured_peak": round(gbs / world / peak, 3)}
        return out
    order = ["sdot", "saxpy", "snrm2", "sasum", "isamax", "sscal", "scopy", "sswap", "srot", "srotm", "sdsdot"]
    routines = table(calls, order, ng, lambda r: BYTES_PER_ELEM[r])
    # srotm's other constructors (Single.hs:475-476 FLAGNEG1: four products; :481-482 FLAG1), FLAG0 is the row above
    neg1 = lb.ModGivensRot(-1, 1.0, 1.0, 1.0, 0.9, -0.3, 0.2, 1.1)
    flag1 = lb.ModGivensRot(1, 1.0, 1.0, 1.0, 0.8, 0.0, 0.0, 0.7)
    routines.update(table({"srotm_FLAGNEG1": lambda: ip.srotm_(neg1, ng, x, 1, y, 1), "srotm_FLAG1": lambda: ip.srotm_(flag1, ng, x, 1, y, 1)},
                          ["srotm_FLAGNEG1", "srotm_FLAG1"], ng, lambda r: 16))

    # ---- the reference's own input distribution (test/Gen.hs:39-44 genNiceFloat: 10% zeros, the rest +-genEveryday, i.e.
    # uniform on (2^-24, 2^24-1), test/Bench.hs:20-31) at the same n: the value-dependent kernels (snrm2's range classes,
    # isamax's running maximum) must not care ----
    routines_everyday = {}
    if not args.no_everyday:
        rng = np.random.default_rng(99 + rank)
        m = min(1 << 24, n)
        kind = rng.random(m)
        tile = rng.uniform(2.0 ** -24, 2.0 ** 24 - 1, m).astype(np.float32)
        tile = np.where(kind < 1 / 9, np.float32(0), np.where(kind < 5 / 9, tile, -tile)).astype(np.float32)
        rev = tile[::-1].copy()
        for v in (x, y):
            for k in range(n // m):
                ctx.wrap(v.ptr + 4 * k * m, m, keepalive=v).upload(tile if v is x else rev)
        times = [time_calls(ctx, barrier, lambda i, f=calls[r]: f(), KR) for r in ("sdot", "snrm2", "sasum", "isamax")]
        for r, ms in zip(("sdot", "snrm2", "sasum", "isamax"), max_over_ranks(times)):
            gbs = BYTES_PER_ELEM[r] * ng / ms / 1e6
            routines_everyday[r] = {"ms": round(ms, 4), "GBps": round(gbs, 1), "pct_of_8TBs_per_gpu": round(100 * gbs / world / NOMINAL_HBM_GBS, 1)}
        routines_everyday["distribution"] = "genNiceFloat: 1/9 zeros, 4/9 +everyday, 4/9 -everyday, everyday = U(2^-24, 2^24-1) (Gen.hs:39-44, 208-214)"

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
            dtimes.append(time_calls(ctx, barrier, lambda i, f=f: f(), KR))
        dtimes = max_over_ranks(dtimes)
        for (r, _), ms in zip(dcalls.items(), dtimes):
            gbs = 2 * BYTES_PER_ELEM["s" + r[1:] if r != "idamax" else "isamax"] * ngd / ms / 1e6
            routines_f64[r] = {"ms": round(ms, 4), "GBps": round(gbs, 1), "pct_of_8TBs_per_gpu": round(100 * gbs / world / NOMINAL_HBM_GBS, 1)}
        del xd, yd, resd

    # ---- strong scaling: the metric's n is the GLOBAL n ----
    # n_global = 2^28 (the metric's size) and 2^30 (BASELINE.json configs[2]) block-sharded over the ranks.  When one rank's
    # operand is <= 2x the 126 MB L2, SETS buffer sets are rotated so that no call finds its operands cached.  The same-box
    # one-GPU time of the same n (rank 0 alone, a context without a communicator, the other ranks idle) gives the speed-up.
    strong = None
    STRONG_R = ("sdot", "saxpy", "snrm2", "sasum", "isamax")
    if world > 1 and not args.no_strong:
        strong = {"routines": list(STRONG_R), "exchange": exchange}
        for lg in (28, 30):
            N = 1 << lg
            per_rank_bytes = 4 * N // world
            SETS = 4 if per_rank_bytes <= 2 * 126_000_000 else (2 if per_rank_bytes <= 4 * 126_000_000 else 1)
            xs_ = [ctx.alloc_sharded(N).fill_hash(31 + i) for i in range(SETS)]
            ys_ = [ctx.alloc_sharded(N).fill_hash(41 + i) for i in range(SETS)]
            cm = [make_calls(xs_[i], ys_[i], N, res) for i in range(SETS)]
            KS = 40 if lg == 28 else 20
            tN = max_over_ranks([time_calls(ctx, barrier, lambda i, r=r: cm[i % SETS][r](), KS, warm=2 * SETS) for r in STRONG_R])
            del cm, xs_, ys_
            # one GPU, same n, same box
            t1 = [0.0] * len(STRONG_R)
            if rank == 0:
                solo = lb.Context(local)
                solo.set_traversal(0)                # like the sharded runs it is compared with: HBM streaming, no L2 walk
                S1 = 2 if 4 * N <= 4 * 126_000_000 else 1
                sx = [solo.alloc(N).fill_hash(31 + i) for i in range(S1)]
                sy = [solo.alloc(N).fill_hash(41 + i) for i in range(S1)]
                sres = solo.alloc(8)
                scm = [make_calls(sx[i], sy[i], N, sres) for i in range(S1)]

                def solo_barrier():
                    solo.sync()
                t1 = [time_calls(solo, solo_barrier, lambda i, r=r: scm[i % S1][r](), KS, warm=2) for r in STRONG_R]
                del scm, sx, sy, sres
                solo.close()
            dist.barrier(group=idle)
            t1 = max_over_ranks(t1)
            strong[f"n_global=2^{lg}"] = {
                r: {"ms": round(a, 4), "ms_1gpu": round(b, 4), "speedup_vs_1gpu": round(b / a, 2),
                    "GBps": round(BYTES_PER_ELEM[r] * N / a / 1e6, 1), "pct_of_8TBs_per_gpu": round(100 * BYTES_PER_ELEM[r] * N / a / 1e6 / world / NOMINAL_HBM_GBS, 1)}
                for r, a, b in zip(STRONG_R, tN, t1)}
            strong[f"n_global=2^{lg}"]["buffer_sets_rotated"] = SETS
        strong["limiter"] = ("at n_global = 2^28 a rank streams 2^28/N elements in tens of microseconds, so the fixed per-call cost shows: the in-kernel "
                             "record exchange (one NVLink round trip, ~3-5 us), the skew between the ranks' independent launches, and the kernel's own "
                             "ramp-up / tail (t0 ~2.5 us on one GPU); at 2^30 the same costs are ~4x smaller relative to the streaming time")

    # ---- the library's default walk (adaptive: from the end the previous kernel left in the L2) on the same back-to-back launches ----
    ctx.set_traversal(1)
    walk_order = ["saxpy", "sscal", "scopy", "sswap", "srot"]
    routines_l2_walk = {r: {"ms": v["ms"], "GBps": v["GBps"], "pct_of_8TBs_per_gpu": v["pct_of_8TBs_per_gpu"]}
                        for r, v in table(calls, walk_order, ng, lambda r: BYTES_PER_ELEM[r]).items()}
    routines_l2_walk["note"] = ("lbk_set_traversal(1), the default: every other launch of a back-to-back chain walks the vectors backwards and finds "
                                "the tail of its predecessor's stream in the L2; algorithmic bytes / time, so part of these bytes never left the L2")

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

    def e2e_step(src_x, src_y):
        for k in range(C):
            xs[k].upload_async(src_x[k * m:(k + 1) * m]); ys[k].upload_async(src_y[k * m:(k + 1) * m])
            ip.saxpy_into(m, SAXPY_A, xs[k], 1, ys[k], 1, ws[k])
            evs[k].record()
            ctx_dl.wait_event(evs[k])
            wd[k].download_async(hy_out[k * m:(k + 1) * m])
        ip.sdot_async(ng, x, 1, y, 1, res)
        ip.snrm2_async(ng, x, 1, res.view(2, 2))
        res.download_async(hres)
        ctx.sync(); ctx_dl.sync()                   # the caller holds the results here

    def time_e2e(src_x, src_y, steps):
        for _ in range(2):
            e2e_step(src_x, src_y)
        barrier()
        tw0 = time.perf_counter()
        for _ in range(steps):
            e2e_step(src_x, src_y)
        barrier()
        return max_over_ranks([(time.perf_counter() - tw0) / steps * 1e3])[0]
    e2e_ms = time_e2e(hx, hy, Ke)
    e2e = {"value": round(STEP_BYTES_PER_ELEM * ng / e2e_ms / 1e6, 2), "unit": "GB/s", "ms_per_step": round(e2e_ms, 3), "steps": Ke,
           "passes_per_step": 1,
           "h2d_bytes_per_step": 8 * n * world, "d2h_bytes_per_step": (4 * n + 32) * world, "chunks": C,
           "timing": "host wall clock around Ke steps, each ending in a synchronise of both streams; max over ranks",
           "api": ("pinned host x,y -> per chunk: lbk_vec_upload_async x2, lbk_saxpy_oop, event -> second context: lbk_wait_event, "
                   "lbk_vec_download_async(y') -> lbk_sdot_async, lbk_snrm2_async, download scalars -> lbk_sync x2")}
    # the same step from memory the caller did NOT get from lbk_host_alloc: pageable (what a plain Data.Vector.Storable /
    # numpy buffer is; the driver stages it through its own bounce buffer, synchronously), and the same buffers after
    # lbk_host_register (page-locked in place)
    px, py = np.array(hx), np.array(hy)
    pageable_ms = time_e2e(px, py, 3)
    with lb.registered(px), lb.registered(py):
        registered_ms = time_e2e(px, py, 3)
    # the reference's own call sequence, nothing hand-pipelined: toDevice x, toDevice y (pageable host vectors), pure saxpy, sdot,
    # snrm2, fromDevice of the result -- synchronous calls of lambda_blas_b200.single, as a Haskell caller of Device.hs would write them
    pz = np.empty(n, dtype=np.float32); pz.fill(0)

    def reference_shaped():
        dx, dy = ctx.to_device(px), ctx.to_device(py)
        z = lb.saxpy(n, SAXPY_A, dx, 1, dy, 1)
        lb.sdot(n, dx, 1, dy, 1); lb.snrm2(n, dx, 1)
        z.download(pz)
    reference_shaped()
    barrier()
    tw0 = time.perf_counter()
    for _ in range(3):
        reference_shaped()
    barrier()
    shaped_ms = max_over_ranks([(time.perf_counter() - tw0) / 3 * 1e3])[0]
    e2e["reference_call_sequence_pageable"] = {
        "value": round(STEP_BYTES_PER_ELEM * ng / shaped_ms / 1e6, 2), "ms_per_step": round(shaped_ms, 3),
        "calls": "to_device(x), to_device(y) [lbk_vec_upload, staged], saxpy (pure), sdot, snrm2, download(z) [lbk_vec_download, staged]; per rank on its own block"}
    del pz
    e2e["from_pageable_host_memory"] = {"value": round(STEP_BYTES_PER_ELEM * ng / pageable_ms / 1e6, 2), "ms_per_step": round(pageable_ms, 3)}
    e2e["from_registered_host_memory"] = {"value": round(STEP_BYTES_PER_ELEM * ng / registered_ms / 1e6, 2), "ms_per_step": round(registered_ms, 3),
                                          "how": "lbk_host_register on the caller's own buffers (not timed: a one-off per buffer)"}

    # ---- parity spot check of the timed configuration (size-independent property) ----
    # after the e2e step: y' = y + a*x exactly (separately rounded), checked on a slice against numpy fp32
    sl = slice(0, 1 << 16)
    want = (hy[sl] + np.float32(SAXPY_A) * hx[sl]).astype(np.float32)
    assert np.array_equal(hy_out[sl].view(np.uint32), want.view(np.uint32)), "saxpy result differs from the fp32 definition"

    # ---- transfers alone: toDevice of one 2^28-element vector (Device.hs toDevice = lbk_vec_upload) by kind of host memory, and
    # every rank's H2D / D2H running together -- the ceiling the end-to-end figure sits under ----
    def wall_ms(f, reps=3):
        f(); barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            f()
        barrier()
        return max_over_ranks([(time.perf_counter() - t0) / reps * 1e3])[0]
    gb = 4 * n * world / 1e6
    up_pinned = wall_ms(lambda: (x.upload_async(hx), ctx.sync()))
    up_pageable = wall_ms(lambda: x.upload(px))                        # lbk_vec_upload: 4 pinned staging lanes inside the library
    up_pageable_driver = wall_ms(lambda: (x.upload_async(px), ctx.sync()))      # cudaMemcpyAsync from pageable memory: the driver's own staging
    down_pageable = wall_ms(lambda: x.download(py))                    # into an existing (already touched) pageable array
    with lb.registered(px):
        up_registered = wall_ms(lambda: (x.upload_async(px), ctx.sync()))
    t0 = time.perf_counter(); lb._ffi.check(lb._ffi.lib().lbk_host_register(px.ctypes.data, px.nbytes)); reg_ms = (time.perf_counter() - t0) * 1e3
    lb._ffi.check(lb._ffi.lib().lbk_host_unregister(px.ctypes.data))
    down_pinned = wall_ms(lambda: (x.download_async(hy_out), ctx.sync()))
    ydl = ctx_dl.wrap(y.ptr, n, keepalive=y)
    both = wall_ms(lambda: (x.upload_async(hx), ydl.download_async(hy_out), ctx.sync(), ctx_dl.sync()))
    del ydl
    # the e2e step's own transfers with nothing else: x and y up (8 B/elem) while y' comes down (4 B/elem) on the second stream
    youtd = ctx_dl.wrap(yout.ptr, n, keepalive=yout)
    floor_ms = wall_ms(lambda: (x.upload_async(hx), y.upload_async(hy), youtd.download_async(hy_out), ctx.sync(), ctx_dl.sync()))
    del youtd
    e2e["transfers_only_ms_per_step"] = round(floor_ms, 3)
    e2e["frac_of_transfers_only"] = round(floor_ms / e2e_ms, 3)
    e2e["bound"] = ("host<->device transfers: the same 8 B/elem up + 4 B/elem down with NO kernels take transfers_only_ms_per_step "
                    "(all ranks at once); compute is hidden under them")
    host_copy = wall_ms(lambda: np.copyto(py, px))
    transfers = {"unit": "GB/s (aggregate over ranks, all ranks transferring at once)", "n_per_gpu": n,
                 "toDevice_pinned": round(gb / up_pinned, 1), "toDevice_pageable": round(gb / up_pageable, 1),
                 "toDevice_pageable_plain_cudaMemcpy": round(gb / up_pageable_driver, 1), "fromDevice_pageable": round(gb / down_pageable, 1),
                 "toDevice_registered": round(gb / up_registered, 1), "host_register_ms_per_GiB": round(reg_ms / (px.nbytes / 2 ** 30), 1),
                 "fromDevice_pinned": round(gb / down_pinned, 1), "h2d_and_d2h_together": round(2 * gb / both, 1),
                 "host_memcpy_1_thread_per_rank": round(2 * gb / host_copy, 1),
                 "note": ("every figure is the aggregate over all ranks transferring at once -- at N > 1 this is the host's ceiling (one VM, one NUMA "
                          "node, shared host DRAM and PCIe root complexes: profiles/r02_host_topology_*.txt); host_memcpy is numpy copyto on every "
                          "rank at once (read + write bytes)")}
    del px, py
    ctx_dl.sync()
    t_load1 = time.time()
    clocks = sampler.stop(t_load0, t_load1) if sampler else None

    del xs, ys, ws, wd, evs, yout
    ctx_dl.close()

    # ---- untimed parity block on this very job (sharded path vs oracle / exact properties) ----
    parity = None
    if not args.no_parity:
        parity = parity_block(lb, np, ctx, world, rank, agree, 30 if world > 1 else 28)

    # ---- the same step from ONE host thread: lbk_init_multi over all GPUs of the job (rank 0 alone; the others wait on the CPU) ----
    del x, y
    single = None
    if not args.no_single:
        if rank == 0:
            single = single_process_arm(lb, np, args, world, local, K, W, P)
        if dist is not None:
            dist.barrier(group=idle)

    if rank == 0:
        saxpy_gbs_gpu = BYTES_PER_ELEM["saxpy"] * n / saxpy_ms_in_step / 1e6
        roofline = {"bound": "hbm", "kernel": f"lbk::{ncu_traffic('saxpy_kernel') or 'map_unit_kernel<AxpyF<float>, 8, 2, 6, 0, 0>'} (saxpy, the largest share of the step)",
                    "achieved": round(saxpy_gbs_gpu, 1), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "frac": round(saxpy_gbs_gpu / peak, 4), "frac_of_nominal_8TBs": round(saxpy_gbs_gpu / NOMINAL_HBM_GBS, 4),
                    "algorithmic_bytes_per_launch": BYTES_PER_ELEM["saxpy"] * n, "launch_ms": round(saxpy_ms_in_step, 4),
                    "launch_ms_how": f"CUDA events around each launch, on the library's stream, in an instrumented repeat of {KI} passes of the timed step",
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
            "metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world), "exchange": exchange,
            "pct_of_8TBs_per_gpu": round(100 * value / world / NOMINAL_HBM_GBS, 1),
            "frac_of_measured_peak_per_gpu": round(value / world / peak, 3),
            "timed_region_ms": round(ms_step * K, 2),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "pdl_launches": int(pdl_launches), "clocks": clocks, "parity_check": parity["status"] if parity else None, "parity": parity,
            "strong": strong, "single_process": single, "transfers": transfers,
            "routines": routines, "routines_l2_walk": routines_l2_walk, "routines_everyday": routines_everyday, "routines_f64": routines_f64,
            "step_routines_ms": {r: round(v, 4) for r, v in per_routine_in_step.items()},
            "last_snrm2": float(step_check[0]),
        }
        emit(line)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def single_process_arm(lb, np, args, world: int, local: int, K: int, W: int, P: int) -> dict:
    """The same step through the single-process multi-device context: ONE host thread, rank r = GPU r.  On one GPU the context
    lists that GPU twice (two ranks, two streams, half the vector each), which is what the driver's one-GPU boxes exercise."""
    from lambda_blas_b200 import inplace as ip
    devs = list(range(world)) if world > 1 else [local, local]
    R = len(devs)
    n = 1 << args.log2n
    ng = n * world
    out = {"devices": devs, "n_global": ng}
    m = lb.MultiContext(devs)
    try:
        x, y = m.alloc_sharded(ng).fill_hash(1), m.alloc_sharded(ng).fill_hash(2)
        res = m.alloc(8)

        def step():
            for _ in range(P):
                ip.sdot_async(ng, x, 1, y, 1, res); ip.saxpy_(ng, SAXPY_A, x, 1, y, 1); ip.snrm2_async(ng, x, 1, res)
        modes = (("threads", True), ("serial", False)) if world > 1 else (("serial", False),)
        for name, threaded in modes:
            m.set_threads(threaded)
            for _ in range(W):
                step()
            m.sync()
            e0, e1 = m.event(), m.event()
            t0 = time.perf_counter()
            e0.record()
            for _ in range(K):
                step()
            e1.record()
            m.sync()
            wall = (time.perf_counter() - t0) * 1e3 / K
            dev = e0.elapsed_ms(e1) / K
            # latency of a synchronous call (host blocks for the scalar): where the launch skew between the ranks shows
            xs, ys = m.alloc_sharded(1 << 22).fill_hash(3), m.alloc_sharded(1 << 22).fill_hash(4)
            for _ in range(20):
                lb.sdot(1 << 22, xs, 1, ys, 1)
            t0 = time.perf_counter()
            for _ in range(200):
                lb.sdot(1 << 22, xs, 1, ys, 1)
            lat = (time.perf_counter() - t0) / 200 * 1e6
            del xs, ys
            out[name] = {"value": round(STEP_BYTES_PER_ELEM * ng * P / dev / 1e6, 1), "unit": "GB/s", "ms_per_step": round(dev, 4),
                         "ms_per_step_host_wall": round(wall, 4), "sdot_2^22_sync_call_us": round(lat, 1)}
        out["last_snrm2"] = float(res.to_host()[0])
        out["timing"] = "CUDA events per device on its own stream, the longest rank counts; host wall clock around the same K steps beside it"
        out["launches"] = m.launch_count()
        del x, y, res
    finally:
        m.close()
    return out



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
    ap.add_argument("--passes", type=int, default=12, help="passes of sdot+saxpy+snrm2 per timed step (12 passes = 10.7 ms per step on one B200)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling section (multi-rank runs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed parity block")
    ap.add_argument("--no-single", action="store_true", help="skip the single-process (lbk_init_multi) arm")
    ap.add_argument("--no-everyday", action="store_true", help="skip the pass over the reference's own input distribution")
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
