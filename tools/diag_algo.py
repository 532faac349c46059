"""Per-layer fwd+bwd time of ConvCheb under the four evaluation-order combinations (DSW_OPT_FWD_ALGO /
DSW_OPT_BWD_ALGO: 1 = TERMS, 2 = CLENSHAW) against the automatic choice, over the cfg3 U-Net layer shapes.

    python tools/diag_algo.py
"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from deepsphere_weather_b200 import _lib  # noqa: E402
from deepsphere_weather_b200 import graphs as G  # noqa: E402
from deepsphere_weather_b200 import layers as L  # noqa: E402


def timed(fn, flush, iters=5, warm=2):
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
    dev = torch.device("cuda:0")
    lib = _lib.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B, K = 32, 4
    laps = {}
    tot = {}
    for idx, (V, Fin, Fout) in enumerate(bench.conv_layer_shapes(32)):
        ns = int(round((V / 12) ** 0.5))
        if ns not in laps:
            laps[ns] = G.healpix_laplacian(ns)
        layer = L.ConvCheb(Fin, Fout, K, laps[ns]).to(dev)
        x = torch.randn(B, V, Fin, device=dev, requires_grad=(idx != 0))
        dy = torch.randn(B, V, Fout, device=dev)

        def fb():
            layer.zero_grad(set_to_none=True)
            if x.requires_grad:
                x.grad = None
            layer(x).backward(dy)

        row = {}
        for fa, ba in [(0, 0), (1, 1), (1, 2), (2, 1), (2, 2)]:
            lib.dsw_set_option(4, fa)
            lib.dsw_set_option(5, ba)
            row[(fa, ba)] = timed(fb, flush)
        lib.dsw_set_option(4, 0)
        lib.dsw_set_option(5, 0)
        auto = (lib.dsw_cheb_fwd_algo(Fin + (-Fin) % 4, Fout + (-Fout) % 4, K), lib.dsw_cheb_bwd_algo(Fin + (-Fin) % 4, Fout + (-Fout) % 4, K))
        best = min((k for k in row if k != (0, 0)), key=lambda k: row[k])
        for k, v in row.items():
            tot[k] = tot.get(k, 0.0) + v
        tot["best"] = tot.get("best", 0.0) + row[best]
        print(f"{V:6d} {Fin:4d}->{Fout:<4d} auto{auto}={row[(0, 0)]:7.1f}  TT={row[(1, 1)]:7.1f} TC={row[(1, 2)]:7.1f} "
              f"CT={row[(2, 1)]:7.1f} CC={row[(2, 2)]:7.1f}  best={best} {row[best]:7.1f}", flush=True)
    print("totals:", {str(k): round(v, 1) for k, v in tot.items()})


if __name__ == "__main__":
    main()
