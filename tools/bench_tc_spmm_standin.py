"""What would a dense-tile tcgen05 (split-bf16) SpMM cost?  

One hop of the metric shape (nside 64: 49 152 nodes, B 32, F 64) as dense tiles: a 128-row tile of the HEALPix k-NN-20
Laplacian gathers ~240 source rows, i.e. out[128 x N] = Ltile[128 x 256] . X[256 x N] — 8 % dense, 12 x the useful
flops, three bf16 MMAs per fp32-accurate product.  That is exactly the MMA / conversion / shared-memory work of this
library's split-bf16 channel mix on a [R x 256] -> 128 problem with R = (V / 128 tiles) * B * F / 128 * 128 rows,
so the existing tcgen05 mix kernel is timed on that shape as a stand-in (it streams its A operand from HBM instead of
gathering it from L2, which is charitable to neither side: 805 MB of A traffic = 125 us at the HBM roofline).

    python tools/bench_tc_spmm_standin.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from deepsphere_weather_b200 import functional as F_  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    timed = bench._event_timer(dev)
    hbm = bench._peaks()[0]
    V, B, F, K = 49152, 32, 64, 4
    tiles = V // 128
    alg = 4.0 * B * V * F * K  # algorithmic bytes of the K - 1 = 3 hop stage
    for ksrc, ncols in [(256, 128), (256, 64), (192, 128)]:
        # rows of the stand-in: every (tile, sample) pair is a [F x ksrc] . [ksrc x 128] product when the tile is the B operand
        rows = tiles * B * F * 128 // ncols
        x = torch.randn(1, rows, ksrc, device=dev)
        w = torch.randn(ncols, ksrc, device=dev)
        with torch.no_grad():
            t = timed(lambda: F_.NodeLinearFunction.apply(x, w, None))
        flops = 2.0 * rows * ksrc * ncols
        print(f"dense-tile hop stand-in: [{rows} x {ksrc}] -> {ncols}: {t * 1e6:7.1f} us per hop, {3 * flops / t / 1e12:6.0f} bf16 TFLOP/s; "
              f"3 hops = {3 * t * 1e6:7.1f} us -> {alg / (3 * t) / 1e9:6.0f} GB/s algorithmic = {alg / (3 * t) / 1e9 / hbm:.2f} of HBM")


if __name__ == "__main__":
    main()
