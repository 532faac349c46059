"""A/B timing of the Chebyshev SpMM stage (dsw_cheb_terms) under the hop tuning options: items per CTA,
small-F routing.  Prints the hop kernel's phase counters for each geometry.

    python tools/diag_hop.py
"""
import ctypes
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepsphere_weather_b200 import _lib  # noqa: E402
from deepsphere_weather_b200 import functional as F_  # noqa: E402
from deepsphere_weather_b200 import graphs as G  # noqa: E402

OPT_DEBUG, OPT_IPC, OPT_SMALL_F, OPT_ROWS = 2, 8, 9, 10


def timed(fn, flush, iters=8, warm=3):
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
    plans = {}
    for nside, B, F in [(32, 32, 64), (32, 32, 128), (16, 32, 256), (8, 32, 512), (64, 32, 64), (32, 32, 24), (32, 32, 4)]:
        if nside not in plans:
            plans[nside] = F_.plan_for(G.healpix_laplacian(nside).to(dev))
        plan = plans[nside]
        V = 12 * nside * nside
        x = torch.randn(B, V, F, device=dev)
        K = 4
        row = []
        for ipc in [0]:
            lib.dsw_set_option(OPT_IPC, ipc)
            row.append((ipc, timed(lambda: F_.cheb_terms(x, plan, K), flush)))
        lib.dsw_set_option(OPT_IPC, 0)
        extra = ""
        for rows in ():
            lib.dsw_set_option(OPT_ROWS, rows)
            extra += f"  rows+{rows}={timed(lambda: F_.cheb_terms(x, plan, K), flush):7.1f}"
        lib.dsw_set_option(OPT_ROWS, 0)
        for teams in (3, 4):
            lib.dsw_set_option(13, teams)
            extra += f"  teams{teams}={timed(lambda: F_.cheb_terms(x, plan, K), flush):7.1f}"
        lib.dsw_set_option(13, 0)
        lib.dsw_set_option(14, 2)
        extra += f"  no-prefetch={timed(lambda: F_.cheb_terms(x, plan, K), flush):7.1f}"
        lib.dsw_set_option(14, 0)
        lib.dsw_set_option(15, 1)
        extra += f"  no-PDL={timed(lambda: F_.cheb_terms(x, plan, K), flush):7.1f}"
        lib.dsw_set_option(15, 0)
        lib.dsw_set_option(11, 8)
        extra += f"  8lanes={timed(lambda: F_.cheb_terms(x, plan, K), flush):7.1f}"
        lib.dsw_set_option(11, 0)
        if F <= 24:
            lib.dsw_set_option(OPT_SMALL_F, 32)
            extra = f"  csr-path {timed(lambda: F_.cheb_terms(x, plan, K), flush):7.1f}"
            lib.dsw_set_option(OPT_SMALL_F, 0)
        print(f"terms nside{nside} B{B} F{F}: " + " ".join(f"ipc{i}={t:7.1f}" for i, t in row) + extra, flush=True)
        buf = (ctypes.c_uint64 * 8)()
        lib.dsw_set_option(OPT_DEBUG, 4)
        F_.cheb_terms(x, plan, K)
        torch.cuda.synchronize()
        lib.dsw_debug_counters(buf, 1)
        F_.cheb_terms(x, plan, K)
        torch.cuda.synchronize()
        lib.dsw_debug_counters(buf, 1)
        lib.dsw_set_option(OPT_DEBUG, 0)
        n = max(buf[4], 1)
        print("   phase cycles per team-item: issue %.0f  wait(tile+Z) %.0f  loop %.0f  store+sync %.0f  items %d" %
              (buf[0] / n, buf[1] / n, buf[2] / n, buf[3] / n, buf[4]), flush=True)
        del x


if __name__ == "__main__":
    main()
