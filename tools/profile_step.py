"""Small drivers for ncu captures (run under gpurun; numbers printed under ncu are never bench values).

    python tools/profile_step.py unet   [steps]      # UNet cfg3 fwd+bwd steps
    python tools/profile_step.py terms  [iters]      # SpMM recurrence, nside 64, B 32, F 64, K 4
    python tools/profile_step.py conv   [iters]      # ConvCheb cfg2 fwd+bwd (nside 32, B 32, 64->64, K 4)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from deepsphere_weather_b200 import functional as F_  # noqa: E402
from deepsphere_weather_b200 import graphs as G  # noqa: E402
from deepsphere_weather_b200 import layers as L  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "unet"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    from deepsphere_weather_b200 import _lib
    lib = _lib.load()
    lib.dsw_set_option(0, int(os.environ.get("DSW_HOP", "0")))
    lib.dsw_set_option(1, int(os.environ.get("DSW_CHUNK", "0")))
    if what == "unet":
        model, V = bench.build_model(dev)
        x = torch.randn(bench.BATCH_PER_GPU, 3, V, 7, device=dev)
        y = torch.randn(bench.BATCH_PER_GPU, 1, V, 2, device=dev)
        crit = torch.nn.MSELoss()
        for _ in range(n):
            loss = crit(model(x), y)
            model.zero_grad(set_to_none=True)
            loss.backward()
        torch.cuda.synchronize()
    elif what == "terms":
        lap = G.healpix_laplacian(64).to(dev)
        plan = F_.plan_for(lap)
        x = torch.randn(32, lap.shape[0], 64, device=dev)
        for _ in range(n):
            F_.cheb_terms(x, plan, 4)
        torch.cuda.synchronize()
    elif what == "conv":
        lap = G.healpix_laplacian(32)
        layer = L.ConvCheb(64, 64, 4, lap).to(dev)
        x = torch.randn(32, 12288, 64, device=dev, requires_grad=True)
        for _ in range(n):
            y = layer(x)
            y.backward(torch.ones_like(y))
        torch.cuda.synchronize()
    elif what == "layer":  # one U-Net layer shape: profile_step.py layer n nside Fin Fout
        nside, Fin, Fout = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
        lap = G.healpix_laplacian(nside)
        layer = L.ConvCheb(Fin, Fout, 4, lap).to(dev)
        x = torch.randn(32, lap.shape[0], Fin, device=dev, requires_grad=True)
        for _ in range(n):
            y = layer(x)
            y.backward(torch.ones_like(y))
        torch.cuda.synchronize()
    print("done", what, n)


if __name__ == "__main__":
    main()
