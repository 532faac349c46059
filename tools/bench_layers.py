"""Per-layer timing of the ConvCheb entry points over the 11 layer shapes of the cfg3 U-Net
(nside 32 -> 16 -> 8, B 32, K 4): the SpMM recurrence alone, the whole forward, the input gradient
and the weight gradient, each with CUDA events and an L2 flush between iterations.  Prints the time,
the algorithmic bytes / flops of SURVEY.md section 8d and the resulting GB/s and TFLOP/s.

    python tools/bench_layers.py [nside] [B] [K]
"""
import ctypes as C
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from deepsphere_weather_b200 import _lib  # noqa: E402
from deepsphere_weather_b200 import functional as F_  # noqa: E402
from deepsphere_weather_b200 import graphs as G  # noqa: E402


def timed(fn, flush, iters=6, warm=2):
    ts = []
    for i in range(warm + iters):
        flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def main():
    nside = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    dev = torch.device("cuda:0")
    lib = _lib.load()
    for k in (0, 1, 2, 3):
        v = os.environ.get(f"DSW_OPT{k}")
        if v:
            lib.dsw_set_option(k, int(v))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    plans = {}
    tot = {"terms": 0.0, "fwd": 0.0, "bwd_data": 0.0, "bwd_weight": 0.0}
    print(f"{'layer':>22s} {'terms us':>9s} {'fwd us':>9s} {'mix us':>8s} {'mix GB/s':>9s} {'mix TF/s':>9s} "
          f"{'bwdD us':>9s} {'bwdW us':>9s} {'wgrad GB/s':>10s}")
    for (V, Fin, Fout) in bench.conv_layer_shapes(nside):
        ns = int(round((V / 12) ** 0.5))
        if ns not in plans:
            plans[ns] = F_.plan_for(G.healpix_laplacian(ns).to(dev))
        plan = plans[ns]
        st = torch.cuda.current_stream().cuda_stream
        x = torch.randn(B, V, Fin, device=dev)
        w = torch.randn(Fin, K, Fout, device=dev) * 0.05
        b = torch.randn(Fout, device=dev)
        dy = torch.randn(B, V, Fout, device=dev)
        y = torch.empty(B, V, Fout, device=dev)
        dx = torch.empty(B, V, Fin, device=dev)
        dw = torch.empty_like(w)
        db = torch.empty_like(b)
        ws_f = torch.empty(lib.dsw_cheb_fwd_workspace_bytes(B, V, Fin, Fout, K), dtype=torch.uint8, device=dev)
        ws_d = torch.empty(lib.dsw_cheb_bwd_data_workspace_bytes(B, V, Fin, Fout, K), dtype=torch.uint8, device=dev)
        ws_w = torch.empty(lib.dsw_cheb_bwd_weight_workspace_bytes(B, V, Fin, Fout, K), dtype=torch.uint8, device=dev)
        terms = torch.empty(max(K - 1, 1), B, V, Fin, device=dev)

        def f_terms():
            _lib.check(lib.dsw_cheb_terms(plan.handle, x.data_ptr(), V * Fin, Fin, terms.data_ptr(), B, Fin, K, st), "terms")

        def f_fwd():
            _lib.check(lib.dsw_cheb_fwd(plan.handle, x.data_ptr(), V * Fin, Fin, w.data_ptr(), b.data_ptr(), y.data_ptr(),
                                        B, Fin, Fout, K, 0, ws_f.data_ptr(), ws_f.numel(), st), "fwd")

        def f_bd():
            _lib.check(lib.dsw_cheb_bwd_data(plan.handle, dy.data_ptr(), w.data_ptr(), dx.data_ptr(), B, Fin, Fout, K,
                                             ws_d.data_ptr(), ws_d.numel(), st), "bwd_data")

        def f_bw():  # saved-terms path: the wgrad kernel + reduction alone
            _lib.check(lib.dsw_cheb_bwd_weight(plan.handle, x.data_ptr(), V * Fin, Fin, dy.data_ptr(), terms.data_ptr(),
                                               dw.data_ptr(), db.data_ptr(), B, Fin, Fout, K, ws_w.data_ptr(),
                                               ws_w.numel(), st), "bwd_weight")

        t_terms, t_fwd, t_bd, t_bw = (timed(f, flush) for f in (f_terms, f_fwd, f_bd, f_bw))
        N = B * V
        mix_us = t_fwd - t_terms
        mix_bytes = 4 * N * (K * Fin + Fout)
        mix_flops = 2 * N * K * Fin * Fout
        print(f"{V:6d} {Fin:4d}->{Fout:<4d}      {t_terms:9.1f} {t_fwd:9.1f} {mix_us:8.1f} {mix_bytes / mix_us / 1e3:9.1f} "
              f"{mix_flops / mix_us / 1e6:9.1f} {t_bd:9.1f} {t_bw:9.1f} {mix_bytes / t_bw / 1e3:10.1f}", flush=True)
        tot["terms"] += t_terms
        tot["fwd"] += t_fwd
        tot["bwd_data"] += t_bd
        tot["bwd_weight"] += t_bw
        del x, w, dy, y, dx, ws_f, ws_d, ws_w, terms
    print("totals (us):", {k: round(v, 1) for k, v in tot.items()}, " sum", round(sum(tot.values()) - tot["terms"], 1))


if __name__ == "__main__":
    main()
