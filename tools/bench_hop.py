"""A/B timing of the Chebyshev SpMM stage (dsw_cheb_terms) and ConvCheb fwd / bwd under the library's
tuning options.  Run under gpurun; prints one line per configuration.

    python tools/bench_hop.py [nside] [B] [F] [K]
"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepsphere_weather_b200 import _lib  # noqa: E402
from deepsphere_weather_b200 import functional as F_  # noqa: E402
from deepsphere_weather_b200 import graphs as G  # noqa: E402
from deepsphere_weather_b200 import layers as L  # noqa: E402

OPT_HOP, OPT_CHUNK = 0, 1


def timed(fn, flush, iters=10, warm=3):
    ts = []
    for i in range(warm + iters):
        if flush is not None:
            flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts), min(ts)


def main():
    nside = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    F = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    K = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    dev = torch.device("cuda:0")
    lib = _lib.load()
    torch.manual_seed(0)
    lap = G.healpix_laplacian(nside).to(dev)
    plan = F_.plan_for(lap)
    V = lap.shape[0]
    x = torch.randn(B, V, F, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    alg = 4 * B * V * F * K + plan.operand_bytes

    lib.dsw_set_option(OPT_HOP, 1)
    lib.dsw_set_option(OPT_CHUNK, 1)
    ref = F_.cheb_terms(x, plan, K)
    configs = [("row-block-L1", 1, 1), ("team", 0, 1)]
    if os.environ.get("DSW_DEBUG_SKIP"):
        configs += [("team-skip-staging", 0, 1), ("team-skip-compute", 0, 1), ("team-2teams", 0, 1)]
    for name, hop, chunk in configs:
        lib.dsw_set_option(OPT_HOP, hop)
        lib.dsw_set_option(OPT_CHUNK, chunk)
        lib.dsw_set_option(2, 1 if name == "team-skip-staging" else 2 if name == "team-skip-compute" else 3 if name == "team-2teams" else 0)
        out = F_.cheb_terms(x, plan, K)
        err = (out - ref).abs().max().item() / ref.abs().max().item()
        med, best = timed(lambda: F_.cheb_terms(x, plan, K), flush)
        print(f"terms nside{nside} B{B} F{F} K{K} {name:22s} median {med:8.1f} us  best {best:8.1f} us  "
              f"{alg / med / 1e3:7.1f} GB/s alg  frac {alg / med / 1e3 / 6545.3:.3f}  maxrel-vs-rb {err:.2e}", flush=True)

    if os.environ.get("DSW_DEBUG_SKIP"):
        import ctypes
        buf = (ctypes.c_uint64 * 8)()
        lib.dsw_set_option(OPT_HOP, 0)
        lib.dsw_set_option(OPT_CHUNK, 1)
        lib.dsw_set_option(2, 4)
        F_.cheb_terms(x, plan, K)
        torch.cuda.synchronize()
        lib.dsw_debug_counters(buf, 1)
        F_.cheb_terms(x, plan, K)
        torch.cuda.synchronize()
        lib.dsw_debug_counters(buf, 1)
        n = max(buf[4], 1)
        print("phase cycles per item (avg over teams): issue %.0f  wait(tile+Z) %.0f  loop %.0f  store+sync %.0f  items %d" %
              (buf[0] / n, buf[1] / n, buf[2] / n, buf[3] / n, buf[4]), flush=True)
        lib.dsw_set_option(2, 0)

    # ConvCheb fwd / fwd+bwd on the same graph (Fin = Fout = F)
    layer = L.ConvCheb(F, F, K, lap).to(dev)
    xg = x.clone().requires_grad_(True)
    lib.dsw_set_option(2, 0)
    for name, hop, chunk in [("row-block-L1", 1, 1), ("team", 0, 1)]:
        lib.dsw_set_option(OPT_HOP, hop)
        lib.dsw_set_option(OPT_CHUNK, chunk)
        with torch.no_grad():
            med, best = timed(lambda: layer(x), flush)
        print(f"conv fwd     {name:22s} median {med:8.1f} us best {best:8.1f}", flush=True)

        def fb():
            y = layer(xg)
            y.backward(y)

        med, best = timed(fb, flush, iters=6)
        print(f"conv fwd+bwd {name:22s} median {med:8.1f} us best {best:8.1f}", flush=True)
    lib.dsw_set_option(OPT_HOP, 0)
    lib.dsw_set_option(OPT_CHUNK, 0)


if __name__ == "__main__":
    main()
