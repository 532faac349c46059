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
        return (float(p["hbm_gbs"]), float(p.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)",
                float(p.get("bf16_tflops_sustained", 1400.0)))
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)", 1400.0


def _profile_traffic(kernel_substr: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel_substr`, from the newest committed
    `ncu --set full` summary under profiles/ (tools/ncu_summary.py output); None when there is none."""
    pdir = os.path.join(ROOT, "profiles")
    best = None
    for name in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if not name.endswith("_ncu.txt"):
            continue
        per_launch, cur, is_kernel = [], {}, False
        with open(os.path.join(pdir, name)) as f:
            for ln in f:
                if ln.startswith("## launch"):
                    if is_kernel and len(cur) == 2:
                        per_launch.append(cur["r"] + cur["w"])
                    cur, is_kernel = {}, kernel_substr in ln
                elif is_kernel and ln.startswith(("dram__bytes_read.sum", "dram__bytes_write.sum")):
                    parts = ln.split()
                    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(parts[2], 1.0)
                    cur["r" if "read" in parts[0] else "w"] = float(parts[1]) * scale
        if is_kernel and len(cur) == 2:
            per_launch.append(cur["r"] + cur["w"])
        if per_launch:
            best = (sum(per_launch) / len(per_launch), name)
    return best


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


def cpu_unet_step_time(batch: int, steps: int, warmup: int, budget_s: float = 0.0):
    """Mean step time of the oracle port on all host cores.  `budget_s` > 0 stops the timed steps early once the
    wall-clock budget is spent (the number of steps actually timed is returned)."""
    from oracle.unet_oracle import oracle_backend

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, V = build_model(torch.device("cpu"), backend=oracle_backend())
    t_start = time.perf_counter()
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
            if budget_s > 0 and time.perf_counter() - t_start > budget_s:
                break
    return sum(times) / len(times), cores, float(loss.item()), len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.cpu_batch
    # the full 32-sample step takes several seconds on the host: time as many of the requested steps as fit ~2.5 minutes
    t, cores, _, n_timed = cpu_unet_step_time(batch, max(args.steps, 1), min(args.warmup, 1), budget_s=150.0)
    val = batch / t
    sample = (f"the full {batch}-sample step (same model, nside {NSIDE}, K {KERNEL_SIZE}), oracle port on {cores} host threads; "
              f"mean of {n_timed} timed steps (of {args.steps} requested, 150 s budget) after {min(args.warmup, 1)} warm-up")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": n_timed, "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cfg3 UNetSpherical nside32->16->8 K=4 B={batch} fwd+bwd", "pool": POOL,
                   "step": "forward + MSE + zero_grad + backward (scalability_plot.py:180-207)"},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------


def _event_timer(device, flush_bytes=256 << 20):
    """CUDA-event timer on the current stream with a 256 MB L2 flush between iterations (median, seconds)."""
    flush = torch.empty(flush_bytes, dtype=torch.uint8, device=device)

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

    return timed


def time_kernel_rooflines(device, hbm_gbs, bf16_sustained):
    """Per-kernel timings of the dominant library calls, CUDA events on the launching stream, L2 flushed (256 MB
    write) between iterations.  SpMM stage: nside 64, B 32, F 64, K 4 (the north-star target shape); channel mix:
    the heaviest U-Net mix (256 -> 512 channels on 393 216 rows); cfg2: one ConvCheb layer."""
    from deepsphere_weather_b200 import _lib
    from deepsphere_weather_b200 import functional as F_
    from deepsphere_weather_b200 import graphs as G
    from deepsphere_weather_b200 import layers as L_

    lib = _lib.load()
    out = {}
    timed = _event_timer(device)

    # --- SpMM recurrence stage (ChebConv SpMM GB/s): ONE fused persistent launch for the K-1 hops ---
    nside, B, F, K = 64, 32, 64, 4
    lap = G.healpix_laplacian(nside).to(device)
    plan = F_.plan_for(lap)
    V = lap.shape[0]
    x = torch.randn(B, V, F, device=device)
    l0 = lib.dsw_launch_count()
    F_.cheb_terms(x, plan, K)
    n_launch = int(lib.dsw_launch_count() - l0)
    t = timed(lambda: F_.cheb_terms(x, plan, K))
    alg_bytes = 4 * B * V * F * K + plan.operand_bytes  # x read once + K-1 terms written + the operator
    achieved = alg_bytes / t / 1e9
    kernel = "hop_chain_kernel" if n_launch == 1 else "hop_team_kernel"
    prof = _profile_traffic(kernel)
    out["roofline"] = {
        "kernel": f"{kernel} (Chebyshev SpMM recurrence, {K - 1} hops in {n_launch} launch(es)), dsw_cheb_terms nside64 B32 F64 K4",
        "bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
        "traffic": prof[0] if prof else None, "traffic_source": f"profiles/{prof[1]}" if prof else None,
        "launches": n_launch, "us_per_launch": t / n_launch * 1e6, "algorithmic_bytes_per_launch": alg_bytes / n_launch,
    }
    del x, lap

    # --- channel mix on the tensor cores: the heaviest U-Net mix shape, through the per-node linear entry point ---
    Bm, Vm, Fi, Fo = 32, 12288, 256, 512
    lin = L_.NodeLinear(Fi, Fo).to(device)
    xm = torch.randn(Bm, Vm, Fi, device=device)
    with torch.no_grad():
        tm = timed(lambda: lin(xm))
    flops = 2.0 * Bm * Vm * Fi * Fo
    peak_tc = bf16_sustained / 3.0  # split-bf16: three bf16 MMAs per fp32-accurate product
    out["roofline_mix"] = {
        "kernel": "mix_tma_kernel (tcgen05 split-bf16 channel mix), dsw_linear_fwd 256->512 on 393216 rows",
        "bound": "tensor", "achieved": flops / tm / 1e12, "peak": peak_tc, "unit": "TFLOP/s",
        "frac": flops / tm / 1e12 / peak_tc, "us_per_launch": tm * 1e6,
        "peak_note": "bf16_tflops_sustained / 3 (three bf16 MMAs per fp32-accurate product)", "traffic": None,
    }
    del xm, lin

    # --- cfg2: one ConvCheb layer, nside 32, B 32, 64 -> 64, K 4 (BASELINE.json configs[1]) ---
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
    # the reference's own torch path for the same layer on this same GPU (cuSPARSE SpMM + cuBLAS fp32)
    try:
        from oracle import cheb_oracle as O

        lap_d, w_d, b_d = lap.to(device), layer.weight.detach(), layer.bias.detach()
        with torch.no_grad():
            t_ref = timed(lambda: O.conv_cheb_layer(lap_d, x, w_d, b_d), iters=5, warm=2)
        out["cfg2_convcheb"]["torch_cuda_fwd_us"] = t_ref * 1e6
        out["cfg2_convcheb"]["speedup_vs_torch_cuda_fwd"] = t_ref / t_fwd
    except Exception as exc:  # never lose the line over a baseline
        out["cfg2_convcheb"]["torch_cuda_fwd_us"] = repr(exc)
    return out


def time_cfg5(device, hbm_gbs):
    """BASELINE.json configs[4]: equiangular 400 x 200 (80 000 nodes, row-major), k-NN-20 Laplacian with irregular
    degree near the poles, ConvCheb K = 6, Cin = Cout = 128, B 8: the SpMM recurrence as a fraction of the HBM roofline."""
    from deepsphere_weather_b200 import functional as F_
    from deepsphere_weather_b200 import graphs as G

    B, F, K = 8, 128, 6
    lap = G.equiangular_laplacian(200, 400).to(device)
    plan = F_.plan_for(lap)
    V = lap.shape[0]
    x = torch.randn(B, V, F, device=device)
    timed = _event_timer(device)
    t = timed(lambda: F_.cheb_terms(x, plan, K), iters=5, warm=2)
    alg = 4 * B * V * F * K + plan.operand_bytes
    return {"workload": "equiangular 400x200 (80000 nodes) k-NN-20, SpMM recurrence K6 C128 B8", "terms_us": t * 1e6,
            "achieved_gbs": alg / t / 1e9, "frac_of_hbm_roofline": alg / t / 1e9 / hbm_gbs}


def time_torch_cuda_baseline(device, steps=5, warmup=2):
    """The reference's own arithmetic on this same GPU: the identical U-Net built on the oracle layers (torch.sparse.mm ->
    cuSPARSE, matmul -> cuBLAS fp32, torch pooling) running the cfg3 step at the full batch ("the stronger,
    same-box comparator").  Device-timed with CUDA events."""
    from oracle.unet_oracle import oracle_backend

    model, V = build_model(device, backend=oracle_backend())
    crit = torch.nn.MSELoss()
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(BATCH_PER_GPU, 3, V, 7, generator=g).to(device)
    y = torch.randn(BATCH_PER_GPU, 1, V, 2, generator=g).to(device)

    def step():
        loss = crit(model(x), y)
        model.zero_grad(set_to_none=True)
        loss.backward()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    e1.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / steps
    return {"value": BATCH_PER_GPU / t, "unit": "samples/s", "ms_per_step": t * 1e3, "steps": steps,
            "what": "same U-Net, same step, torch-CUDA path of the reference arithmetic (cuSPARSE + cuBLAS fp32) on this GPU, B 32"}


def time_ar_rollout(device, steps=5):
    """The autoregressive training step of the reference's loop (xforecasting.AutoregressiveTraining with ar_iterations = 2:
    three forward passes with the predictions fed back, WeightedMSELoss per iteration, one backward) on cfg3's model at 8
    samples — the small-batch regime of the 8-GPU run — eager and replayed from one CUDA graph."""
    from deepsphere_weather_b200.ar import ARRollout
    from deepsphere_weather_b200.ddp import FlatGradBucket
    from deepsphere_weather_b200.losses import WeightedMSELoss

    B, T, Fd, Fb, Fs, ar_it = 8, 3, 2, 1, 4, 2
    model, V = build_model(device)
    bucket = FlatGradBucket(model)
    crit = WeightedMSELoss(weights=torch.rand(V, device=device) + 0.5)
    roll = ARRollout(model, crit, ar_it)
    mk = lambda *s: torch.randn(*s, device=device)
    args = (mk(B, T, V, Fd), mk(B, T + ar_it + 1, V, Fb), mk(V, Fs), mk(B, ar_it + 1, V, Fd))

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / steps

    t_eager = timed(lambda: roll.step(*args, zero_grad=bucket.zero_))
    res = {"workload": f"AR rollout: cfg3 model, B {B}, ar_iterations {ar_it} (3 forward passes), WeightedMSELoss, one backward",
           "eager_ms_per_step": t_eager * 1e3, "eager_samples_per_s": B / t_eager}
    try:
        roll.capture(*args, zero_grad=bucket.zero_)
        t_graph = timed(lambda: roll.replay(*args))
        res.update({"cuda_graph_ms_per_step": t_graph * 1e3, "cuda_graph_samples_per_s": B / t_graph})
    except Exception as exc:
        res["cuda_graph_error"] = repr(exc)[:200]
        torch.cuda.synchronize()
    return res


def time_cfg4_strong(device, rank, world, steps, lib):
    """BASELINE.json configs[3] / north star: UNetSpherical nside 64 (49 152 nodes), K 4, GLOBAL batch 64 sharded over the
    ranks (64 / 32 / 16 / 8 samples per GPU at N = 1 / 2 / 4 / 8): strong scaling.  Forward + backward are replayed from a
    CUDA graph (one capture per process; ~600 kernel launches per step otherwise bound the small-batch step on the host);
    the gradient all-reduce follows the replay.  Returns a dict for rank 0 (max over ranks, device-timed)."""
    import torch.distributed as dist

    from deepsphere_weather_b200.ddp import FlatGradBucket

    global_batch = 64
    if global_batch % world:
        return {"skipped": f"global batch {global_batch} does not divide over {world} ranks"}
    Bl = global_batch // world
    torch.manual_seed(4321 + rank)
    model, V = build_model(device, nside=64)
    bucket = FlatGradBucket(model)
    crit = torch.nn.MSELoss()
    x = torch.randn(Bl, 3, V, 7, device=device)
    y = torch.randn(Bl, 1, V, 2, device=device)

    def fwd_bwd():
        loss = crit(model(x), y)
        bucket.zero_()
        loss.backward()
        return loss

    for _ in range(3):
        fwd_bwd()
        bucket.allreduce_mean()
    torch.cuda.synchronize()
    l0 = lib.dsw_launch_count()
    fwd_bwd()
    launches = int(lib.dsw_launch_count() - l0)
    graph, graph_err = None, None
    if os.environ.get("DSW_BENCH_NO_GRAPH", "0") != "1":
        try:
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                fwd_bwd()
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                fwd_bwd()
            graph.replay()
            torch.cuda.synchronize()
        except Exception as exc:
            graph, graph_err = None, repr(exc)[:200]
            torch.cuda.synchronize()

    def step():
        if graph is not None:
            graph.replay()
        else:
            fwd_bwd()
        bucket.allreduce_mean()

    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        tt = torch.tensor([secs], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        secs = float(tt.item())
    res = {
        "workload": "cfg4 UNetSpherical nside64->32->16 (49152 nodes) K=4, global batch 64, fwd+bwd (+ grad all-reduce)",
        "scaling": "strong", "global_batch": global_batch, "batch_per_gpu": Bl, "n_gpus": world, "steps": steps,
        "ms_per_step": secs / steps * 1e3, "samples_per_s": global_batch * steps / secs,
        "cuda_graph": graph is not None, "library_launches_per_step": launches,
    }
    if graph_err:
        res["cuda_graph_error"] = graph_err
    del graph, model, bucket, x, y
    torch.cuda.empty_cache()
    return res


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
    hbm_gbs, bf16_tflops, peak_src, bf16_sustained = _peaks()

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

    # ---- cfg4: the north star's strong-scaling configuration, at every N (all ranks take part) ----
    cfg4 = None
    if not args.no_cfg4:
        del bufs
        torch.cuda.empty_cache()
        try:
            cfg4 = time_cfg4_strong(device, rank, world, max(3, min(args.steps, 10)), lib)
        except Exception as exc:  # never lose the headline line
            cfg4 = {"error": repr(exc)[:300]}

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
    if cfg4 is not None:
        line["cfg4_strong"] = cfg4
    if world == 1:
        del model, bucket
        torch.cuda.empty_cache()
        try:
            line.update(time_kernel_rooflines(device, hbm_gbs, bf16_sustained))
            line["nodes_channels_per_s"] = line["cfg2_convcheb"]["nodes_channels_per_s_fwd"]
        except Exception as exc:  # never lose the headline line
            line["roofline"] = {"error": repr(exc)}
        try:
            line["ar_rollout"] = time_ar_rollout(device)
        except Exception as exc:
            line["ar_rollout"] = {"error": repr(exc)[:200]}
        try:
            line["cfg5_equiangular"] = time_cfg5(device, hbm_gbs)
        except Exception as exc:
            line["cfg5_equiangular"] = {"error": repr(exc)[:200]}
        try:
            line["torch_cuda_baseline"] = time_torch_cuda_baseline(device)
            line["torch_cuda_baseline"]["speedup_of_value"] = value / line["torch_cuda_baseline"]["value"]
        except Exception as exc:
            line["torch_cuda_baseline"] = {"error": repr(exc)[:200]}
        if not args.no_cpu_baseline:
            t, cores, _, _n = cpu_unet_step_time(args.cpu_batch, 1, 1)
            line["cpu_baseline"] = {
                "value": args.cpu_batch / t, "unit": "samples/s", "cores": cores, "kind": "port",
                "sample": f"the full {args.cpu_batch}-sample step, 1 warm-up + 1 timed step, oracle port (torch CPU, {cores} threads)",
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
    ap.add_argument("--cpu-batch", type=int, default=BATCH_PER_GPU, help="batch of the CPU arm's step (default: the full step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the cfg4 strong-scaling leg (nside 64, global batch 64)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
