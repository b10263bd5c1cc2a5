#!/usr/bin/env python
"""bench.py — the headline measurement: Float64 evaluations/s of a 3-D order-(32,32,32) vector-valued
(SVector{3}) ChebPoly (BASELINE.json configs[2]) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode val|jac] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch of B synthetic points per GPU (weak scaling: B is
fixed per GPU; the north_star's 10^9 points are B x steps x GPUs).  Prints ONE JSON line on rank 0.

  value     device-resident inputs, K steps timed with CUDA events on the launching stream, max over ranks
  e2e       same metric through the public host API (fastcheb.ChebPoly.__call__ -> fastcheb_eval, MEM_HOST)
            with pinned HOST buffers: H2D of the points and D2H of the results are inside the timed region
  roofline  algorithmic FP64 flops (W_val = 2*K*S per point) / kernel time vs the MEASURED
            FP64 tensor-core peak of this pool (profiles/r01_micro_fp64.jsonl, dmma884: 37.0 TFLOP/s)
  cpu_baseline  the oracle's OpenMP C restatement of the reference on this box's host cores, bounded sample
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = (33, 33, 33)
K = 3
LB = np.array([-1.0, -1.0, -1.0])
UB = np.array([1.0, 1.0, 1.0])
COEF_SEED, PTS_SEED = 4, 5


def work_per_point(dims, K, cplx=1, jac=False):
    """SURVEY.md §8d canonical algorithmic flops per evaluated point."""
    N = len(dims)
    Sd = [int(np.prod(dims[d:])) for d in range(N)]
    if jac:
        return 2 * K * cplx * sum((d + 2) * Sd[d] for d in range(N))
    return 2 * K * cplx * sum(Sd)


def fp64_peak_tflops():
    """Measured FP64 peak of this pool's B200 (bench_micro/micro_fp64.cu, committed results)."""
    path = os.path.join(ROOT, "profiles", "r01_micro_fp64.jsonl")
    best = 0.0
    try:
        for line in open(path):
            r = json.loads(line)
            if r.get("test", "").startswith("dmma884") or r.get("test", "").startswith("dfma_ch"):
                best = max(best, float(r.get("tflops", 0.0)))
    except OSError:
        pass
    if best > 0:
        return best, "measured: profiles/r01_micro_fp64.jsonl (DMMA m8n8k4 register-resident loop, B200 @1965 MHz)"
    return 37.2, "fallback: nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz"


def coefficients():
    rng = np.random.default_rng(COEF_SEED)
    return rng.uniform(-1.0, 1.0, DIMS + (K,))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples within the upper half of the observed power range
        thr = (max(power) + min(power)) / 2
        load = [s for s, p in zip(sm, power) if p >= thr] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


_REAL_STDOUT = None


def protect_stdout():
    """stdout carries the ONE JSON line.  Native libraries (NCCL's NCCL_DEBUG=VERSION banner is a bare printf) also
    write to fd 1, so fd 1 is pointed at stderr for the lifetime of the process and the JSON line goes to a private
    duplicate of the original stdout.  The driver's NCCL_DEBUG / NCCL_DEBUG_FILE settings are left untouched."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _REAL_STDOUT


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def cpu_reference_rate(mode: str, sample_pts: int, reps: int = 3):
    """The reference's algorithm on the host cores: the oracle's OpenMP C restatement of eval.jl
    (Threads.@threads-over-c(y) analogue).  Returns (points/s, threads)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    so = os.path.join(ROOT, "oracle", "libcheb_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    import cheb_oracle as orc

    orc.set_num_threads()  # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    poly = orc.OracleChebPoly(coefficients(), LB, UB)
    rng = np.random.default_rng(PTS_SEED)
    pts = rng.uniform(-1.0, 1.0, (sample_pts, len(DIMS)))
    # value mode: all K components of an element per pass (the SIMD the reference gets from SVector{3}); bit-identical
    # to the component-by-component restatement (tests/test_oracle_golden.py)
    f = poly.jacobian if mode == "jac" else poly.eval_simd
    f(pts[:1000])
    best = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        f(pts)
        best = max(best, sample_pts / (time.perf_counter() - t0))
    return best, orc.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_sample
    t_all = []
    for i in range(args.warmup):
        cpu_reference_rate(args.mode, min(sample, 20000), reps=1)
    rate, threads = 0.0, 1
    for i in range(args.steps):
        t0 = time.perf_counter()
        r, threads = cpu_reference_rate(args.mode, sample, reps=1)
        t_all.append(time.perf_counter() - t0)
        rate = max(rate, r)
    W = work_per_point(DIMS, K, jac=args.mode == "jac")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": rate, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(t_all)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.mode), "sample_points_per_step": sample},
        "cpu_baseline": {"value": rate, "unit": "points/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} points per step, OpenMP over points, SIMD over components, oracle/cheb_oracle.c "
                                   "(Julia is not installable here; this is the line-faithful C restatement)"},
        "e2e": {"value": rate, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gflops_algorithmic": rate * W / 1e9,
    }
    emit(line)


def measured_traffic(key, units_key, units):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r01_traffic.json), when
    the capture was taken at this launch size; else None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[key]
        return int(t["dram_bytes"]) if int(t[units_key]) == int(units) else None
    except Exception:
        return None


def hbm_peak_gbs():
    """Measured HBM copy bandwidth of this pool (driver-written MEASURED_PEAKS.json), else the profiling guide's fallback."""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured: MEASURED_PEAKS.json hbm_gbs (torch copy, read+write)"
    except Exception:
        return 6400.0, "fallback: B200_PROFILING.md measured-copy figure"


def cpu_fit_rate(n, howmany, dt):
    """The reference's fit on the host cores: FFTW.r2r(REDFT00) stand-in = scipy/pocketfft dct(type=1) with all
    workers (libfftw3 is not in this image) + the normalisation passes of interp.jl:49-54.  Returns (fits/s, threads)."""
    import scipy.fft as sf
    rng = np.random.default_rng(8)
    vals = rng.uniform(-1, 1, (howmany, n)).astype(dt)
    th = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    best = 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        c = sf.dct(vals, type=1, axis=1, workers=th)
        c /= 2 * (n - 1)
        c[:, 1:n - 1] *= 2
        best = max(best, howmany / (time.perf_counter() - t0))
    return best, th


def run_fit(args):
    """configs[4]: batches of independent 1-D order-64 fits (chebcoefs, interp.jl:39-57) — the HBM-bound leg.
    A step = one batched DCT-I fit of `--howmany` lines per GPU; fits are sharded per GPU (no collective)."""
    n = 65
    dt = np.float64 if args.dtype == "f64" else np.float32
    rs = np.dtype(dt).itemsize
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank != 0:
            return
        hm = min(args.howmany, 200_000)
        r = 0.0
        for _ in range(max(1, args.steps)):
            rr, th = cpu_fit_rate(n, hm, dt)
            r = max(r, rr)
        emit(({"impl": "reference", "metric": "1-D order-64 DCT-I fits/sec", "value": r, "unit": "fits/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                          "config": {"workload": f"configs[4]: {hm} independent 1-D order-64 fits per step"},
                          "cpu_baseline": {"value": r, "unit": "fits/s", "cores": th, "kind": "port",
                                           "sample": f"{hm} fits, scipy/pocketfft dct(type=1, workers={th}) + normalisation"},
                          "e2e": {"value": r, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    import torch
    import torch.distributed as dist
    from fastcheb import _capi
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        protect_stdout()   # NCCL's banner / log lines must not land in the JSON stream; its env is not touched
        dist.init_process_group("nccl", device_id=dev)
    L = _capi.lib()
    M = int(args.howmany)
    tdt = torch.float64 if dt == np.float64 else torch.float32
    bytes_per_fit = 2 * rs * n
    # rotate over enough buffer pairs that consecutive steps never find their lines in the 126 MB L2
    nbuf = max(2, int(np.ceil(300e6 / (M * bytes_per_fit))) + 1)
    g = torch.Generator(device=dev)
    g.manual_seed(8 + rank)
    vals = [torch.rand((M, n), dtype=tdt, device=dev, generator=g) * 2 - 1 for _ in range(nbuf)]
    outs = [torch.empty((M, n), dtype=tdt, device=dev) for _ in range(nbuf)]
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    dims = _capi.i64_array((n,))
    code = _capi.F64 if dt == np.float64 else _capi.F32

    def launch(i):
        # device-resident, asynchronous: enqueue only (the finite-check flag is read back by the synchronous API)
        rc = L.fastcheb_fit_async(1, dims, code, 0, 1, M, vals[i % nbuf].data_ptr(), outs[i % nbuf].data_ptr(), None, None,
                                  local, sp)
        if rc != 0:
            raise RuntimeError(_capi.last_error())

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        launch(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    l0 = L.fastcheb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        launch(i)
    e1.record(stream)
    barrier()
    launches = L.fastcheb_launch_count() - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = M * world * args.steps / (ms_max * 1e-3)
    kern_ms = ms / args.steps
    peak, peak_src = hbm_peak_gbs()
    achieved = M * bytes_per_fit / (kern_ms * 1e-3) / 1e9

    # parity spot check against the oracle on the first lines (checker only)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cheb_oracle as orc
    got = outs[0][:64].cpu().numpy().astype(np.float64)
    want = np.stack([orc.chebcoefs(v, 1) for v in vals[0][:64].cpu().numpy().astype(np.float64)])
    err = float(np.abs(got - want).max() / np.abs(want).max())

    e2e = None
    if not args.no_e2e:
        Me = min(M, 1 << 20)
        hv = torch.empty((Me, n), dtype=tdt).pin_memory()
        hv.copy_(vals[0][:Me].cpu())
        hc = torch.empty((Me, n), dtype=tdt).pin_memory()
        hm = torch.empty(Me, dtype=torch.float64).pin_memory()

        def e2e_step():
            rc = L.fastcheb_fit(1, dims, code, 0, 1, Me, hv.data_ptr(), hc.data_ptr(), hm.data_ptr(), _capi.MEM_HOST, local, None)
            if rc != 0:
                raise RuntimeError(_capi.last_error())
            return float(hc[0, 0])
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        nst = max(3, min(args.steps, 10))
        for _ in range(nst):
            e2e_step()
        dtw = time.perf_counter() - t0
        tt = torch.tensor([dtw], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": Me * world * nst / float(tt.item()), "unit": "fits/s", "h2d_bytes_per_step": Me * n * rs,
               "d2h_bytes_per_step": Me * n * rs + Me * 8, "fits_per_step_per_gpu": Me,
               "timer": "host wall clock around the synchronous C-ABI call (finite check + per-fit max|c| included)"}
    if rank == 0:
        cpu = None
        if not args.no_cpu:
            r, th = cpu_fit_rate(n, min(M, 200_000), dt)
            cpu = {"value": r, "unit": "fits/s", "cores": th, "kind": "port",
                   "sample": f"{min(M, 200_000)} fits, best of 3, scipy/pocketfft dct(type=1, workers={th}) + normalisation "
                             "passes (stand-in for the reference's FFTW.r2r REDFT00; libfftw3 is not in this image)"}
        emit(({
            "metric": "1-D order-64 DCT-I fits/sec", "value": value, "unit": "fits/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"configs[4]: {M} independent 1-D order-64 fits per GPU per step (chebcoefs)",
                       "l2": f"{nbuf} rotating input/output buffer pairs of {M * bytes_per_fit / 1e6:.0f} MB: no step finds its lines in L2",
                       "parallelism": f"fits sharded over {world} GPU(s), no collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic("fit1d_" + args.dtype, "launch_fits", M), "bytes_per_fit": bytes_per_fit, "kernel_ms": kern_ms, "peak_source": peak_src},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "max_rel_err_vs_oracle": err}))
    if distributed:
        dist.destroy_process_group()


WORKLOAD = "eval3d"
METRIC = "Float64 evals/sec (3D order-32 ChebPoly)"


def select_workload(name):
    """eval3d = BASELINE configs[2] (the headline); eval2d = configs[1]: 2-D order (64,64) real Float64, value or
    value + chebgradient — the same bench contract on the second evaluation config."""
    global WORKLOAD, METRIC, DIMS, K, LB, UB
    WORKLOAD = name
    if name == "eval2d":
        DIMS, K = (65, 65), 1
        LB, UB = np.array([-1.0, -1.0]), np.array([1.0, 1.0])
        METRIC = "Float64 evals/sec (2D order-64 ChebPoly)"


def workload_name(mode):
    if WORKLOAD == "eval2d":
        what = "value c.(ys)" if mode == "val" else "value + chebgradient"
        return f"configs[1]: 2-D order (64,64) real Float64 ChebPoly, {what}, uniform random points"
    what = "value c.(ys)" if mode == "val" else "value + chebjacobian"
    return f"configs[2]: 3-D order (32,32,32) SVector{{3}} Float64 ChebPoly, {what}, uniform random points"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="val", choices=["val", "jac"])
    ap.add_argument("--batch", type=int, default=1 << 24, help="points per GPU per step")
    ap.add_argument("--algo", default="auto", choices=["auto", "clenshaw", "dmma"])
    ap.add_argument("--cpu-sample", type=int, default=200_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="only the headline measurement (no `legs` object)")
    ap.add_argument("--quick-legs", action="store_true", help="legs with 2 timed repetitions instead of 5")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"], help="N>1: how values reach rank 0")
    ap.add_argument("--workload", default="eval3d", choices=["eval3d", "eval2d", "fit1d"],
                    help="eval3d = the headline (configs[2]); eval2d = configs[1] (2-D order 64); "
                         "fit1d = configs[4] batched DCT-I fits (HBM-bound leg)")
    ap.add_argument("--howmany", type=int, default=1 << 20, help="fit1d: fits per GPU per step")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="fit1d: element type")
    args = ap.parse_args()
    if args.workload == "fit1d":
        return run_fit(args)
    select_workload(args.workload)
    if args.workload == "eval2d" and args.cpu_sample == 200_000:
        args.cpu_sample = 5_000_000  # a 2-D point is 26x cheaper than a headline point: keep the CPU sample in seconds
    if args.warmup < 3:
        args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import fastcheb
    from fastcheb import _capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        protect_stdout()   # NCCL's banner / log lines must not land in the JSON stream; its env is not touched
        dist.init_process_group("nccl", device_id=dev)

    L = _capi.lib()
    jac = args.mode == "jac"
    B = int(args.batch)
    N, E = len(DIMS), K

    # ---- the polynomial: rank 0 owns the coefficients, NCCL broadcasts them (north_star: eval sharding)
    if distributed:
        from fastcheb import sharded
        carr, lb_, ub_ = sharded.broadcast_coefficients(coefficients() if rank == 0 else None, LB if rank == 0 else None,
                                                        UB if rank == 0 else None, src=0)
    else:
        carr, lb_, ub_ = coefficients(), LB, UB
    poly = fastcheb.ChebPoly(carr, lb_, ub_, device=local)
    poly.set_algo({"auto": fastcheb.ALGO_AUTO, "clenshaw": fastcheb.ALGO_CLENSHAW, "dmma": fastcheb.ALGO_DMMA}[args.algo])
    h = poly._handle(np.float64)
    algo_used = {1: "clenshaw", 2: "dmma"}[poly.get_algo(jac)]

    # ---- synthetic points, generated on the device (inputs resident in HBM before the timed region)
    g = torch.Generator(device=dev)
    g.manual_seed(PTS_SEED + 1000 * rank)
    pts = torch.rand((B, N), dtype=torch.float64, device=dev, generator=g) * 2.0 - 1.0
    out_v = torch.empty((B, E), dtype=torch.float64, device=dev)
    out_J = torch.empty((B, N, E), dtype=torch.float64, device=dev)
    status = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def measure(jac_mode: bool, steps: int, warmup: int, sample_clocks: bool):
        """K steps of the hot path with device-resident inputs, CUDA events on the launching stream, max over ranks.
        N > 1: the results of all ranks land in ONE buffer on rank 0 inside the timed region.  Default: fused eval +
        gather — every rank's kernel stores straight into rank 0's HBM through a CUDA-IPC mapping (NVLink peer stores,
        fastcheb.sharded.PeerGather; values and, in Jacobian mode, the 96 B/point Jacobians); --gather nccl uses a
        torch.distributed gather after the kernel instead."""
        do_gather = distributed and not args.no_gather
        gathered_v = gathered_J = peer_v = peer_J = None
        ov_ptr, oJ_ptr = out_v.data_ptr(), out_J.data_ptr()
        if do_gather and args.gather == "nccl":
            gathered_v = [torch.empty_like(out_v) for _ in range(world)] if rank == 0 else None
            gathered_J = [torch.empty_like(out_J) for _ in range(world)] if (rank == 0 and jac_mode) else None
        elif do_gather:
            from fastcheb import sharded
            peer_v = sharded.PeerGather(total_rows=B * world, row_reals=E, itemsize=8, dst=0, device=local)
            ov_ptr = peer_v.slice_ptr(rank * B)
            if jac_mode:
                peer_J = sharded.PeerGather(total_rows=B * world, row_reals=E * N, itemsize=8, dst=0, device=local)
                oJ_ptr = peer_J.slice_ptr(rank * B)

        def step():
            if jac_mode:
                rc = L.fastcheb_eval_jac_async(h, pts.data_ptr(), B, ov_ptr, oJ_ptr, status.data_ptr(), sp)
            else:
                rc = L.fastcheb_eval_async(h, pts.data_ptr(), B, ov_ptr, status.data_ptr(), sp)
            if rc != 0:
                raise RuntimeError(_capi.last_error())
            if do_gather and args.gather == "nccl":
                dist.gather(out_v, gathered_v, dst=0)  # results gathered on rank 0 over NVLink (NCCL)
                if jac_mode:
                    dist.gather(out_J, gathered_J, dst=0)

        for _ in range(warmup):
            step()
        barrier()
        if int(status.item()) != 0x7FFFFFFFFFFFFFFF:
            raise RuntimeError("out-of-domain point in the synthetic batch")
        sampler = ClockSampler(local)
        if rank == 0 and sample_clocks:
            sampler.start()
            time.sleep(0.3)
        barrier()
        l0 = L.fastcheb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        barrier()
        launches = L.fastcheb_launch_count() - l0
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        # kernel-only time (local output, no gather) for the roofline: one kernel per step
        if do_gather:
            barrier()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(stream)
            for _ in range(steps):
                if jac_mode:
                    L.fastcheb_eval_jac_async(h, pts.data_ptr(), B, out_v.data_ptr(), out_J.data_ptr(), status.data_ptr(), sp)
                else:
                    L.fastcheb_eval_async(h, pts.data_ptr(), B, out_v.data_ptr(), status.data_ptr(), sp)
            k1.record(stream)
            barrier()
            kern_ms = k0.elapsed_time(k1) / steps
        else:
            kern_ms = ms / steps
        gather_check = None
        if peer_v is not None:
            # every rank's slice of the gathered buffers equals what that rank computes locally
            if jac_mode:
                L.fastcheb_eval_jac_async(h, pts.data_ptr(), B, out_v.data_ptr(), out_J.data_ptr(), status.data_ptr(), sp)
            else:
                L.fastcheb_eval_async(h, pts.data_ptr(), B, out_v.data_ptr(), status.data_ptr(), sp)
            torch.cuda.synchronize(dev)
            mine = (out_v[:4096].cpu().numpy(), out_J[:4096].cpu().numpy() if jac_mode else None)
            lst = [None] * world
            dist.all_gather_object(lst, mine)
            if rank == 0:
                gather_check = max(float(np.abs(peer_v.read(r * B, 4096).reshape(4096, E) - lst[r][0]).max()) for r in range(world))
                if jac_mode:
                    gather_check = max(gather_check, max(float(np.abs(peer_J.read(r * B, 4096).reshape(4096, N, E) - lst[r][1]).max())
                                                         for r in range(world)))
            peer_v.close()
            if peer_J is not None:
                peer_J.close()
        W = work_per_point(DIMS, K, jac=jac_mode)
        return {"value": B * world * steps / (ms_max * 1e-3), "ms_per_step": ms_max / steps, "kern_ms": kern_ms,
                "launches": int(launches), "clocks": clocks, "gather_check": gather_check, "W": W,
                "achieved": B * W / (kern_ms * 1e-3) / 1e12, "gathered": bool(do_gather)}

    main_res = measure(jac, args.steps, args.warmup, True)
    value, ms_max, kern_ms, launches, clocks = (main_res["value"], main_res["ms_per_step"] * args.steps, main_res["kern_ms"],
                                                main_res["launches"], main_res["clocks"])
    W, achieved, gather_check, do_gather = main_res["W"], main_res["achieved"], main_res["gather_check"], main_res["gathered"]
    peak, peak_src = fp64_peak_tflops()

    # The other mode of the same workload, all ranks, gather inside the timed region (configs[2]: "10^9 evals +
    # chebjacobian sharded over 1/2/4/8 B200"): reported under legs.
    other = None
    if not args.no_legs:
        other = measure(not jac, max(3, min(args.steps, 6)), 3, False)

    # ---- e2e: the public host API with HOST buffers (H2D + kernel + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        Be = min(B, 1 << 23)
        hp = torch.empty((Be, N), dtype=torch.float64).pin_memory()
        hp.copy_(pts[:Be].cpu())
        hv = torch.empty((Be, E), dtype=torch.float64).pin_memory()
        hJ = torch.empty((Be, N, E), dtype=torch.float64).pin_memory() if jac else None

        def e2e_step(ip, iv, iJ):
            if jac:
                rc = L.fastcheb_eval_jac(h, ip, Be, iv, iJ, _capi.MEM_HOST, None)
            else:
                rc = L.fastcheb_eval(h, ip, Be, iv, _capi.MEM_HOST, None)
            if rc != 0:
                raise RuntimeError(_capi.last_error())

        def e2e_time(ip, iv, iJ, read):
            for _ in range(2):
                e2e_step(ip, iv, iJ)
            barrier()
            t0 = time.perf_counter()
            nst = max(3, min(args.steps, 10))
            for _ in range(nst):
                e2e_step(ip, iv, iJ)
                read()  # host read of the step's result
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if distributed:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return Be * world * nst / float(tt.item())

        e2e = {"value": e2e_time(hp.data_ptr(), hv.data_ptr(), hJ.data_ptr() if jac else None, lambda: float(hv[0, 0])),
               "unit": "points/s", "h2d_bytes_per_step": Be * N * 8, "d2h_bytes_per_step": Be * E * 8 * (1 + (N if jac else 0)),
               "points_per_step_per_gpu": Be, "buffers": "pinned host memory",
               "timer": "host wall clock around the synchronous C-ABI call, max over ranks"}
        # spot-check e2e output against the device path
        L.fastcheb_eval_async(h, pts.data_ptr(), B, out_v.data_ptr(), status.data_ptr(), sp)
        torch.cuda.synchronize(dev)
        e2e["max_abs_diff_vs_device_path"] = float((hv[:1000] - out_v[:1000].cpu()).abs().max())
        # the same call with PAGEABLE host buffers (plain malloc'ed numpy arrays: what a Julia Vector is)
        pp = np.empty((Be, N), dtype=np.float64)
        pp[...] = hp.numpy()
        pv = np.empty((Be, E), dtype=np.float64)
        pJ = np.empty((Be, N, E), dtype=np.float64) if jac else None
        e2e["pageable"] = {"value": e2e_time(pp.ctypes.data, pv.ctypes.data, pJ.ctypes.data if jac else None, lambda: float(pv[0, 0])),
                           "unit": "points/s", "buffers": "pageable host memory (numpy / malloc)"}
        e2e["pageable"]["max_abs_diff_vs_pinned"] = float(np.abs(pv[:1000] - hv[:1000].numpy()).max())
        del pp, pv, pJ
        # N > 1: the same batch driven from ONE host process through fastcheb_multi_eval (rank 0 drives all N GPUs,
        # the other ranks are idle at the barrier)
        if distributed and not jac:
            barrier()
            # the other ranks wait on the rendezvous store (a CPU-side wait): an NCCL barrier would keep a spinning kernel
            # resident on their GPUs, which are about to be driven by rank 0
            store = dist.distributed_c10d._get_default_store()
            if rank == 0:
                e2e["single_process"] = single_process_e2e(L, _capi, carr, lb_, ub_, world, Be, N, E, hp)
                store.set("fastcheb_single_process_done", "1")
            else:
                store.wait(["fastcheb_single_process_done"])
            barrier()

    if rank == 0:
        cpu = None
        if not args.no_cpu:
            r, th = cpu_reference_rate(args.mode, args.cpu_sample)
            cpu = {"value": r, "unit": "points/s", "cores": th, "kind": "port",
                   "sample": f"{args.cpu_sample} uniform random points of the same workload, best of 3, "
                             "oracle/cheb_oracle.c (OpenMP over points, SIMD over the K components like the reference's SVector arithmetic)"}
        line = {
            "metric": METRIC, "value": value, "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "gpu_launches": int(launches),
            "config": {"workload": workload_name(args.mode), "points_per_gpu_per_step": B,
                       "coef_bytes": int(np.prod(DIMS)) * K * 8, "algo": algo_used,
                       "l2": f"inputs {B * N * 8 / 2**20:.0f} MiB + outputs per step exceed the 126 MB L2",
                       "parallelism": f"points sharded over {world} GPU(s), coefficients NCCL-broadcast, "
                                      + (("results gathered on rank 0 inside the timed region: "
                                          + ("kernel stores straight into rank 0's HBM over NVLink (CUDA-IPC peer buffer, fused eval+gather)"
                                             if args.gather == "peer" else "NCCL gather after the kernel"))
                                         if do_gather else "results stay sharded")},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak,
                         "traffic": measured_traffic(WORKLOAD + ("_jac" if jac else "_val"), "launch_points", B),
                         "traffic_note": "DRAM bytes per launch (ncu, profiles/r01_traffic.json); algorithmic: "
                                         f"{B * 8 * (N + E * (1 + (N if jac else 0)))} B of points in + results out",
                         "flops_per_point": W, "kernel_ms": kern_ms, "peak_source": peak_src,
                         "note": "FP64 tensor-core (DMMA) pipe; algorithmic flops, not executed flops"},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "clocks": clocks,
        }
        if gather_check is not None:
            line["gather_max_abs_diff_vs_local"] = gather_check
        if not args.no_parity:
            line["parity"] = parity_record(fastcheb, poly, carr, lb_, ub_)
        legs = {}
        if other is not None:
            nm = WORKLOAD + ("_val" if jac else "_jac")
            legs[nm] = {"value": other["value"], "unit": "points/s", "ms_per_step": other["ms_per_step"], "n_gpus": world,
                        "gathered_on_rank0": other["gathered"],
                        "roofline": {"bound": "tensor", "achieved": other["achieved"], "peak": peak, "unit": "TFLOP/s",
                                     "frac": other["achieved"] / peak, "flops_per_point": other["W"], "kernel_ms": other["kern_ms"]}}
            if other["gather_check"] is not None:
                legs[nm]["gather_max_abs_diff_vs_local"] = other["gather_check"]
        line["legs"] = legs
    del pts, out_v, out_J
    poly.close()
    torch.cuda.empty_cache()
    if rank == 0 and not args.no_legs:
        # every other leg of BASELINE.json's configs, short runs on this rank's GPU, each with its own roofline
        import bench_legs
        line["legs"].update(bench_legs.run_all(local, quick=args.quick_legs))
    if rank == 0:
        emit(line)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def single_process_e2e(L, _capi, carr, lb_, ub_, world, Be, N, E, hp):
    """`c.(ys)` on N GPUs from ONE host process: fastcheb_multi_create + fastcheb_multi_eval with pinned host buffers
    holding world x Be points (the per-rank e2e batch replicated), H2D + kernels + D2H inside the timed region."""
    import torch

    h = ctypes.c_void_p()
    import fastcheb
    flat = np.ascontiguousarray(fastcheb._to_julia_flat(np.asarray(carr), N))
    rc = L.fastcheb_multi_create(ctypes.byref(h), world, None, N, _capi.i64_array(DIMS), _capi.F64, 0, K, flat.ctypes.data,
                                 _capi.MEM_HOST, _capi.dbl_array(lb_), _capi.dbl_array(ub_))
    if rc != 0:
        return {"error": _capi.last_error()}
    tot = Be * world
    big_p = torch.empty((tot, N), dtype=torch.float64).pin_memory()
    for r in range(world):
        big_p[r * Be:(r + 1) * Be].copy_(hp)
    big_v = torch.empty((tot, E), dtype=torch.float64).pin_memory()
    for _ in range(2):
        rc = L.fastcheb_multi_eval(h, big_p.data_ptr(), tot, big_v.data_ptr(), _capi.MEM_HOST)
    t0 = time.perf_counter()
    nst = 3
    for _ in range(nst):
        rc |= L.fastcheb_multi_eval(h, big_p.data_ptr(), tot, big_v.data_ptr(), _capi.MEM_HOST)
        float(big_v[0, 0])
    dt = time.perf_counter() - t0
    same = bool(torch.equal(big_v[:1000], big_v[(world - 1) * Be:(world - 1) * Be + 1000]))
    L.fastcheb_multi_destroy(h)
    return {"value": tot * nst / dt, "unit": "points/s", "n_gpus": world, "points_per_step": tot, "rc": int(rc),
            "replicated_blocks_agree": same, "entry": "fastcheb_multi_eval (one process, one worker thread + stream per GPU)"}


def parity_record(fastcheb, poly, carr, lb_, ub_):
    """Measured parity of the default path against the oracle on a seeded sample (faces and near-face points included):
    worst |dv| / (16 eps sum|c|) and, per gradient entry, worst |dJ_d| over the differentiated-series bound
    16 eps (sum_k |c_k| k_d^2) 2/(ub_d-lb_d) and over the literal 16 eps sum|c| 2/(ub_d-lb_d)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cheb_oracle as orc

    rng = np.random.default_rng(123)
    N = len(DIMS)
    lb, ub = np.asarray(lb_, dtype=float), np.asarray(ub_, dtype=float)
    n = 4096
    p = lb + (ub - lb) * rng.random((n, N))
    p[0], p[1] = lb, ub
    m = n // 8
    d = rng.integers(0, N, m)
    p[np.arange(8, 8 + m), d] = np.where(rng.random(m) < 0.5, lb[d] + (ub - lb)[d] * 1e-4 * rng.random(m), ub[d] - (ub - lb)[d] * 1e-4 * rng.random(m))
    v0, J0 = orc.OracleChebPoly(carr, lb, ub).jacobian(p)
    v, J = fastcheb.chebjacobian(poly, p)
    eps = np.finfo(np.float64).eps
    sumc = float(np.abs(carr).sum())
    wj, sj = [], []
    for dd in range(N):
        k2 = (np.arange(DIMS[dd], dtype=float) ** 2).reshape([-1 if i == dd else 1 for i in range(np.ndim(carr))])
        dj = float(np.abs(np.asarray(J)[..., dd] - np.asarray(J0)[..., dd]).max())
        sc = 2.0 / (ub[dd] - lb[dd])
        wj.append(dj / (16 * eps * float((np.abs(carr) * k2).sum()) * sc))
        sj.append(dj / (16 * eps * sumc * sc))
    return {"points": n, "value_over_tol": float(np.abs(np.asarray(v) - np.asarray(v0)).max() / (16 * eps * sumc)),
            "jac_over_weighted_tol": max(wj), "jac_over_literal_tol": max(sj),
            "tolerances": "value: 16 eps sum|c|; gradient entries: 16 eps sum_k|c_k| k_d^2 * 2/(ub_d-lb_d) (weighted) and "
                          "16 eps sum|c| * 2/(ub_d-lb_d) (literal; the oracle itself is 0.2-176x this bound from the long-double "
                          "Jacobian, profiles/r02_jac_parity.jsonl)"}


if __name__ == "__main__":
    main()
