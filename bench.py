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


def cpu_reference_rate(mode: str, sample_pts: int, reps: int = 3):
    """The reference's algorithm on the host cores: the oracle's OpenMP C restatement of eval.jl
    (Threads.@threads-over-c(y) analogue).  Returns (points/s, threads)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    so = os.path.join(ROOT, "oracle", "libcheb_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    import cheb_oracle as orc

    poly = orc.OracleChebPoly(coefficients(), LB, UB)
    rng = np.random.default_rng(PTS_SEED)
    pts = rng.uniform(-1.0, 1.0, (sample_pts, 3))
    f = poly.jacobian if mode == "jac" else poly.__call__
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
        "metric": "Float64 evals/sec (3D order-32 ChebPoly)", "value": rate, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(t_all)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.mode), "sample_points_per_step": sample},
        "cpu_baseline": {"value": rate, "unit": "points/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} points per step, OpenMP over points, oracle/cheb_oracle.c "
                                   "(Julia is not installable here; this is the line-faithful C restatement)"},
        "e2e": {"value": rate, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gflops_algorithmic": rate * W / 1e9,
    }
    print(json.dumps(line))


def workload_name(mode):
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
    args = ap.parse_args()
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
        dist.init_process_group("nccl", device_id=dev)

    L = _capi.lib()
    jac = args.mode == "jac"
    B = int(args.batch)
    N, E = 3, K

    # ---- the polynomial: rank 0 owns the coefficients, NCCL broadcasts them (north_star: eval sharding)
    ct = torch.empty(DIMS + (K,), dtype=torch.float64, device=dev)
    if rank == 0:
        ct.copy_(torch.from_numpy(coefficients()))
    if distributed:
        dist.broadcast(ct, src=0)
    poly = fastcheb.ChebPoly(ct.cpu().numpy(), LB, UB, device=local)
    poly.set_algo({"auto": fastcheb.ALGO_AUTO, "clenshaw": fastcheb.ALGO_CLENSHAW, "dmma": fastcheb.ALGO_DMMA}[args.algo])
    h = poly._handle(np.float64)
    algo_used = {1: "clenshaw", 2: "dmma"}[poly.get_algo(jac)]

    # ---- synthetic points, generated on the device (inputs resident in HBM before the timed region)
    g = torch.Generator(device=dev)
    g.manual_seed(PTS_SEED + 1000 * rank)
    pts = torch.rand((B, N), dtype=torch.float64, device=dev, generator=g) * 2.0 - 1.0
    out_v = torch.empty((B, E), dtype=torch.float64, device=dev)
    out_J = torch.empty((B, N, E), dtype=torch.float64, device=dev) if jac else None
    status = torch.zeros(1, dtype=torch.int64, device=dev)
    gathered = None
    if distributed and not args.no_gather and not jac:
        gathered = [torch.empty_like(out_v) for _ in range(world)] if rank == 0 else None
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)

    def step():
        if jac:
            rc = L.fastcheb_eval_jac_async(h, pts.data_ptr(), B, out_v.data_ptr(), out_J.data_ptr(), status.data_ptr(), sp)
        else:
            rc = L.fastcheb_eval_async(h, pts.data_ptr(), B, out_v.data_ptr(), status.data_ptr(), sp)
        if rc != 0:
            raise RuntimeError(_capi.last_error())
        if distributed and not args.no_gather and not jac:
            dist.gather(out_v, gathered, dst=0)  # results gathered on rank 0 over NVLink

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    if int(status.item()) != 0x7FFFFFFFFFFFFFFF:
        raise RuntimeError("out-of-domain point in the synthetic batch")

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    l0 = L.fastcheb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    launches = L.fastcheb_launch_count() - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = B * world * args.steps / (ms_max * 1e-3)

    # kernel-only time (no gather) for the roofline: one kernel per step
    if distributed and not args.no_gather and not jac:
        barrier()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        for _ in range(args.steps):
            L.fastcheb_eval_async(h, pts.data_ptr(), B, out_v.data_ptr(), status.data_ptr(), sp)
        k1.record(stream)
        barrier()
        kern_ms = k0.elapsed_time(k1) / args.steps
    else:
        kern_ms = ms / args.steps
    W = work_per_point(DIMS, K, jac=jac)
    peak, peak_src = fp64_peak_tflops()
    achieved = B * W / (kern_ms * 1e-3) / 1e12

    # ---- e2e: the public host API with pinned host buffers (H2D + kernel + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        Be = min(B, 1 << 23)
        hp = torch.empty((Be, N), dtype=torch.float64).pin_memory()
        hp.copy_(pts[:Be].cpu())
        hv = torch.empty((Be, E), dtype=torch.float64).pin_memory()
        hJ = torch.empty((Be, N, E), dtype=torch.float64).pin_memory() if jac else None

        def e2e_step():
            if jac:
                rc = L.fastcheb_eval_jac(h, hp.data_ptr(), Be, hv.data_ptr(), hJ.data_ptr(), _capi.MEM_HOST, None)
            else:
                rc = L.fastcheb_eval(h, hp.data_ptr(), Be, hv.data_ptr(), _capi.MEM_HOST, None)
            if rc != 0:
                raise RuntimeError(_capi.last_error())
            return float(hv[0, 0])  # host read of the step's result

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        nst = max(3, min(args.steps, 10))
        for _ in range(nst):
            e2e_step()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": Be * world * nst / float(tt.item()), "unit": "points/s",
               "h2d_bytes_per_step": Be * N * 8, "d2h_bytes_per_step": Be * E * 8 * (1 + (N if jac else 0)),
               "points_per_step_per_gpu": Be, "timer": "host wall clock around the synchronous C-ABI call, max over ranks"}
        # spot-check e2e output against the device path
        chk = float((hv[:1000] - out_v[:1000].cpu()).abs().max())
        e2e["max_abs_diff_vs_device_path"] = chk

    if rank == 0:
        cpu = None
        if not args.no_cpu:
            r, th = cpu_reference_rate(args.mode, args.cpu_sample)
            cpu = {"value": r, "unit": "points/s", "cores": th, "kind": "port",
                   "sample": f"{args.cpu_sample} uniform random points of the same workload, best of 3, "
                             "oracle/cheb_oracle.c (OpenMP over points)"}
        line = {
            "metric": "Float64 evals/sec (3D order-32 ChebPoly)", "value": value, "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.mode), "points_per_gpu_per_step": B,
                       "coef_bytes": int(np.prod(DIMS)) * K * 8, "algo": algo_used,
                       "l2": f"inputs {B * N * 8 / 2**20:.0f} MiB + outputs per step exceed the 126 MB L2",
                       "parallelism": f"points sharded over {world} GPU(s), coefficients NCCL-broadcast, "
                                      + ("values gathered on rank 0 (NCCL gather) inside the timed region"
                                         if (distributed and not args.no_gather and not jac) else "results stay sharded")},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": None,
                         "flops_per_point": W, "kernel_ms": kern_ms, "peak_source": peak_src,
                         "note": "FP64 tensor-core (DMMA) pipe; algorithmic flops, not executed flops"},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
