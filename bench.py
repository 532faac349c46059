#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (contract: task prompt §"Measurement").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

Workload (BASELINE.json metric, cfg3): UNetSpherical, HEALPix nside 32 (12 288 nodes) -> 16 -> 8,
K = 4, 7 input variables x 3 time steps -> 2 outputs, batch 32 per GPU, fp32.  One *step* is what
the reference's own timing harness runs (scripts_figs/scalability_plot.py:180-207): forward, MSE
loss, zero the gradients, backward — here followed, for N > 1, by the one gradient all-reduce.

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` is the
same metric through the public module API starting from pinned host buffers (H2D of the inputs and
D2H of the loss inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

NSIDE = 32
BATCH_PER_GPU = 32
KERNEL_SIZE = 4
POOL = "interp"  # shipped config: configs/UNetSpherical/Healpix_400km/InterpPool-Graph_knn.json
METRIC = "UNetSpherical fwd+bwd samples/s @ HEALPix nside=32, K=4"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi SM clock / throttle-reason sampler running during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        mhz, mx, reasons = [], None, set()
        for s in self.samples:
            try:
                mhz.append(float(s[0]))
                mx = float(s[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(mhz) if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz)}


def build_model(device, backend=None, nside=NSIDE):
    from deepsphere_weather_b200 import models as M

    V = 12 * nside * nside
    kw = dict(kernel_size_conv=KERNEL_SIZE, pool_method=POOL)
    if backend is not None:
        kw["backend"] = backend
    model = M.UNetSpherical(M.default_tensor_info(V), "healpix", {"subdivisions": nside, "nest": True}, **kw)
    M.deterministic_fill(model, seed=0, rezero=1.0)
    return model.to(device), V


def conv_layer_shapes(nside=NSIDE):
    V0, V1, V2 = 12 * nside**2, 12 * (nside // 2) ** 2, 12 * (nside // 4) ** 2
    return [(V0, 21, 64), (V0, 64, 128), (V1, 128, 192), (V1, 192, 256), (V2, 256, 512), (V2, 512, 256),
            (V1, 512, 256), (V1, 256, 128), (V0, 256, 128), (V0, 128, 64), (V0, 64, 2)]


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ------------------------------------------------------------------------------------------------


def cpu_unet_step_time(batch: int, steps: int, warmup: int):
    from oracle.unet_oracle import oracle_backend

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, V = build_model(torch.device("cpu"), backend=oracle_backend())
    g = torch.Generator().manual_seed(1)
    x = torch.randn(batch, 3, V, 7, generator=g)
    y_obs = torch.randn(batch, 1, V, 2, generator=g)
    crit = torch.nn.MSELoss()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss = crit(model(x), y_obs)
        model.zero_grad(set_to_none=True)
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times), cores, float(loss.item())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.cpu_batch
    t, cores, _ = cpu_unet_step_time(batch, max(args.steps, 1), args.warmup)
    val = batch / t
    sample = f"batch {batch} of the {BATCH_PER_GPU}-sample step (same model, nside {NSIDE}, K {KERNEL_SIZE}); mean of {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg3 UNetSpherical nside32->16->8 K=4 B=32 fwd+bwd", "pool": POOL,
                   "step": "forward + MSE + zero_grad + backward (scalability_plot.py:180-207)"},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------


def time_kernel_rooflines(device, hbm_gbs, bf16_tflops):
    """Per-kernel timings of the two dominant library calls, CUDA events on the launching stream,
    L2 flushed (256 MB write) between iterations.  SpMM stage: nside 64, B 32, F 64, K 4 (the
    north-star target shape); channel mix: the heaviest U-Net layer."""
    from deepsphere_weather_b200 import functional as F_
    from deepsphere_weather_b200 import graphs as G

    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def timed(fn, iters=8, warm=3):
        ts = []
        for i in range(warm + iters):
            flush.fill_(i & 0xFF)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            if i >= warm:
                ts.append(e0.elapsed_time(e1) * 1e-3)
        return statistics.median(ts)

    # --- SpMM recurrence stage (ChebConv SpMM GB/s) ---
    nside, B, F, K = 64, 32, 64, 4
    lap = G.healpix_laplacian(nside).to(device)
    plan = F_.plan_for(lap)
    V = lap.shape[0]
    x = torch.randn(B, V, F, device=device)
    t = timed(lambda: F_.cheb_terms(x, plan, K))
    n_launch = K - 1
    alg_bytes = 4 * B * V * F * K + plan.operand_bytes  # x read + K-1 terms written ... per stage
    per_launch_bytes = alg_bytes / n_launch
    achieved = alg_bytes / t / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the three hop launches in
    # profiles/r01k_hop_team_kernel_ncu.txt (one `ncu --set full` capture of this same call):
    # (462.4 + 363.9) + (867.3 + 376.2) + (867.3 + 374.9) MB over 3 launches.  Each unfused hop moves
    # ~3 planes (gather source, k-2 term, output) where the K-plane accounting counts 4/3.
    ncu_traffic_bytes_per_launch = 1104.0e6
    out["roofline"] = {
        "kernel": "hop_team_kernel (Chebyshev SpMM hop), dsw_cheb_terms nside64 B32 F64 K4", "bound": "hbm",
        "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
        "traffic": ncu_traffic_bytes_per_launch,
        "launches": n_launch, "us_per_launch": t / n_launch * 1e6, "algorithmic_bytes_per_launch": per_launch_bytes,
    }
    del x, lap

    # --- cfg2: one ConvCheb layer, nside 32, B 32, 64 -> 64, K 4 (BASELINE.json configs[1]) ---
    from deepsphere_weather_b200 import layers as L_

    nside, B, F, K = 32, 32, 64, 4
    lap = G.healpix_laplacian(nside)
    layer = L_.ConvCheb(F, F, K, lap).to(device)
    V = lap.shape[0]
    x = torch.randn(B, V, F, device=device)
    with torch.no_grad():
        t_fwd = timed(lambda: layer(x))
    xg = x.clone().requires_grad_(True)
    dy = torch.randn(B, V, F, device=device)

    def fwd_bwd():
        layer.zero_grad(set_to_none=True)
        xg.grad = None
        layer(xg).backward(dy)

    t_fb = timed(fwd_bwd)
    nnz_bytes = F_.plan_for(layer.laplacian).operand_bytes
    fused_bytes = 4 * B * V * F * 2 + nnz_bytes + 4 * K * F * F + 4 * F      # fused-layer compulsory traffic
    flops = 2 * (nnz_bytes // 8) * F * B * (K - 1) + 2 * B * V * K * F * F
    out["cfg2_convcheb"] = {
        "workload": "ConvCheb nside32 (12288 nodes) B32 64->64 K4", "fwd_us": t_fwd * 1e6, "fwd_bwd_us": t_fb * 1e6,
        "nodes_channels_per_s_fwd": B * V * F / t_fwd, "nodes_channels_per_s_fwd_bwd": B * V * F / t_fb,
        "fused_layer_algorithmic_bytes": fused_bytes, "fwd_frac_of_hbm_roofline": fused_bytes / t_fwd / 1e9 / hbm_gbs,
        "fwd_tflops_fp32_equivalent": flops / t_fwd / 1e12,
    }
    return out


def run_ours(args):
    from deepsphere_weather_b200 import _lib
    from deepsphere_weather_b200.ddp import FlatGradBucket

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    import torch.distributed as dist

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()
    hbm_gbs, bf16_tflops, peak_src = _peaks()

    torch.manual_seed(1234 + rank)
    model, V = build_model(device)
    bucket = FlatGradBucket(model)
    crit = torch.nn.MSELoss()
    B = BATCH_PER_GPU
    x_dev = torch.randn(B, 3, V, 7, device=device)
    y_dev = torch.randn(B, 1, V, 2, device=device)
    x_host = x_dev.cpu().pin_memory()
    y_host = y_dev.cpu().pin_memory()
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step(x, y):
        loss = crit(model(x), y)
        bucket.zero_()
        loss.backward()
        bucket.allreduce_mean()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.dsw_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        secs = e0.elapsed_time(e1) * 1e-3
        if world > 1:
            t = torch.tensor([secs], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        return secs, lib.dsw_launch_count() - l0

    # ---- device-resident timing ----
    for _ in range(max(args.warmup, 3)):
        step(x_dev, y_dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    secs, launches = timed_region(lambda: step(x_dev, y_dev), args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: pinned host -> device, public API, loss back to host ----
    # Every step's inputs come from pinned host memory and its loss goes back to the host.  The copy of
    # step i + 1 is issued on a side stream into the other device buffer while step i computes (what a
    # DataLoader with pin_memory / non_blocking does); all copies happen inside the timed region.
    copy_stream = torch.cuda.Stream(device)
    bufs = [(torch.empty_like(x_dev), torch.empty_like(y_dev)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        xb, yb = bufs[i % 2]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])  # the step that last used this buffer is done
            xb.copy_(x_host, non_blocking=True)
            yb.copy_(y_host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_run(steps):
        cur = torch.cuda.current_stream(device)
        for ev in consumed:
            ev.record(cur)
        prefetch(0)
        for i in range(steps):
            if i + 1 < steps:
                prefetch(i + 1)
            cur.wait_event(ready[i % 2])
            xb, yb = bufs[i % 2]
            loss = step(xb, yb)
            consumed[i % 2].record(cur)
            loss_host.copy_(loss.detach(), non_blocking=True)
            cur.synchronize()  # the caller reads the loss every step

    e2e_run(2)
    e2e_secs, _ = timed_region(lambda: e2e_run(args.steps), 1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_samples = B * world * args.steps
    value = total_samples / secs
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"cfg3 UNetSpherical nside32->16->8 K=4 B={B}/GPU fwd+bwd", "pool": POOL,
            "global_batch": B * world, "nodes": V, "parallelism": f"dp{world} (batch shards, NCCL grad all-reduce)",
            "step": "forward + MSE + zero_grad + backward (scalability_plot.py:180-207)" + (" + grad all-reduce" if world > 1 else ""),
            "mix_mode": "tcgen05 split-bf16" if lib.dsw_get_mix_mode() == 1 else "fp32 CUDA cores",
            "l2": "per-step working set (~2.4 GB fwd) >> 126 MB L2; no explicit flush in the step loop",
            "peaks": peak_src,
        },
        "clocks": clocks,
        "e2e": {"value": total_samples / e2e_secs, "unit": "samples/s",
                "h2d_bytes_per_step": (x_host.numel() + y_host.numel()) * 4 * world, "d2h_bytes_per_step": 4 * world},
        "gpu_launches": int(launches),
    }
    if world == 1:
        try:
            line.update(time_kernel_rooflines(device, hbm_gbs, bf16_tflops))
            line["nodes_channels_per_s"] = line["cfg2_convcheb"]["nodes_channels_per_s_fwd"]
        except Exception as exc:  # never lose the headline line
            line["roofline"] = {"error": repr(exc)}
        if not args.no_cpu_baseline:
            t, cores, _ = cpu_unet_step_time(args.cpu_batch, 1, 1)
            line["cpu_baseline"] = {
                "value": args.cpu_batch / t, "unit": "samples/s", "cores": cores, "kind": "port",
                "sample": f"batch {args.cpu_batch} of the {B}-sample step, 1 warm-up + 1 timed step, oracle port (torch CPU)",
            }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--cpu-batch", type=int, default=4, help="bounded sample of the step for the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
