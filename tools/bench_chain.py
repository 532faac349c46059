"""A/B timing of the Chebyshev SpMM stage (dsw_cheb_terms): fused persistent chain kernel vs hop-by-hop launches,
over L2 budgets / pass floors.  Run under gpurun; one line per configuration.

    python tools/bench_chain.py [nside] [B] [F] [K]
"""
import ctypes
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepsphere_weather_b200 import _lib  # noqa: E402
from deepsphere_weather_b200 import functional as F_  # noqa: E402
from deepsphere_weather_b200 import graphs as G  # noqa: E402

OPT_DEBUG, OPT_NO_CHAIN, OPT_L2, OPT_MIN_PASS = 2, 17, 18, 19


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
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    dev = torch.device("cuda:0")
    lib = _lib.load()
    torch.manual_seed(0)
    lap = G.healpix_laplacian(nside).to(dev)
    plan = F_.plan_for(lap)
    V = lap.shape[0]
    x = torch.randn(B, V, F, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    alg = 4 * B * V * F * K + plan.operand_bytes
    per_sample = 3 * V * F * 4

    lib.dsw_set_option(OPT_NO_CHAIN, 1)
    ref = F_.cheb_terms(x, plan, K)
    configs = [("hop-by-hop", 1, 0, 0), ("chain default", 0, 0, 0)]
    for s in (1, 2, 3, 4, 6, 8, 16, 32):
        if s <= B:
            configs.append((f"chain S={s}", 0, s * per_sample, 1))
    for name, no_chain, l2, min_pass in configs:
        lib.dsw_set_option(OPT_NO_CHAIN, no_chain)
        lib.dsw_set_option(OPT_L2, l2)
        lib.dsw_set_option(OPT_MIN_PASS, min_pass)
        out = F_.cheb_terms(x, plan, K)
        same = torch.equal(out, ref)
        med, best = timed(lambda: F_.cheb_terms(x, plan, K), flush)
        print(f"terms nside{nside} B{B} F{F} K{K} {name:16s} median {med:8.1f} us  best {best:8.1f} us  "
              f"{alg / med / 1e3:7.1f} GB/s alg  frac {alg / med / 1e3 / peak:.3f}  bit-identical {same}", flush=True)
    lib.dsw_set_option(OPT_L2, 0)
    lib.dsw_set_option(OPT_MIN_PASS, 0)
    lib.dsw_set_option(OPT_NO_CHAIN, 0)
    if os.environ.get("DSW_CHAIN_DEBUG"):
        for name, dbg in [("claim-ahead", 32), ("relaxed-done(unsafe)", 64), ("relaxed-done+no-Z", 80), ("no-stores", 8), ("no-Z", 16), ("no-stores-no-Z", 24), ("no-loop", 2), ("no-loop-stores-Z", 26)]:
            lib.dsw_set_option(OPT_DEBUG, dbg)
            med, best = timed(lambda: F_.cheb_terms(x, plan, K), flush)
            print(f"debug {name:18s} median {med:8.1f} us  best {best:8.1f} us", flush=True)
            if dbg == 32:
                buf = (ctypes.c_uint64 * 16)()
                lib.dsw_set_option(OPT_DEBUG, 36)
                lib.dsw_debug_chain_counters(buf, 1)
                F_.cheb_terms(x, plan, K)
                torch.cuda.synchronize()
                lib.dsw_debug_chain_counters(buf, 1)
                n, m = max(buf[4], 1), max(buf[15], 1)
                print("  team: wait-next %.0f  zg+transfer-wait %.0f  loop %.0f  stores %.0f | issuer: metadata %.0f  wait-loop-end %.0f  deps+issue %.0f  late %d of %d" %
                      (buf[0] / n, buf[1] / n, buf[2] / n, buf[3] / n, buf[8] / m, buf[9] / m, buf[10] / m, buf[14], buf[15]), flush=True)
        lib.dsw_set_option(OPT_DEBUG, 0)

    if os.environ.get("DSW_CHAIN_PHASES"):
        buf = (ctypes.c_uint64 * 16)()
        lib.dsw_set_option(OPT_DEBUG, 4)
        F_.cheb_terms(x, plan, K)
        torch.cuda.synchronize()
        lib.dsw_debug_chain_counters(buf, 1)
        F_.cheb_terms(x, plan, K)
        torch.cuda.synchronize()
        lib.dsw_debug_chain_counters(buf, 1)
        n = max(buf[4], 1)
        print("team cycles per item: wait-next %.0f  zg+transfer-wait %.0f  loop %.0f  stores %.0f  items %d" %
              (buf[0] / n, buf[1] / n, buf[2] / n, buf[3] / n, buf[4]), flush=True)
        m = max(buf[15], 1)
        print("issuer cycles per item: metadata %.0f  wait-loop-end %.0f  deps+issue %.0f  late %d of %d" %
              (buf[8] / m, buf[9] / m, buf[10] / m, buf[14], buf[15]), flush=True)
        lib.dsw_set_option(OPT_DEBUG, 0)


if __name__ == "__main__":
    main()
