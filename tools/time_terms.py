"""CUDA-event timing of the SpMM recurrence (dsw_cheb_terms) with an L2 flush between iterations, for same-box A/B runs of two
builds (DSW_LIB_PATH=...):  python tools/time_terms.py [nside B F K]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from deepsphere_weather_b200 import functional as F_  # noqa: E402
from deepsphere_weather_b200 import graphs as G  # noqa: E402


def main():
    nside, B, F, K = [int(a) for a in sys.argv[1:5]] if len(sys.argv) > 4 else (64, 32, 64, 4)
    dev = torch.device("cuda:0")
    plan = F_.plan_for(G.healpix_laplacian(nside).to(dev))
    x = torch.randn(B, 12 * nside * nside, F, device=dev)
    timed = bench._event_timer(dev)
    t = timed(lambda: F_.cheb_terms(x, plan, K), iters=12)
    alg = 4.0 * B * x.shape[1] * F * K
    print(f"{os.environ.get('DSW_LIB_PATH', 'libdsw.so')}: terms nside {nside} B {B} F {F} K {K}: {t * 1e6:.1f} us = "
          f"{alg / t / 1e9:.0f} GB/s = {alg / t / 1e9 / bench._peaks()[0]:.3f} of HBM")


if __name__ == "__main__":
    main()
