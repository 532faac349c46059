"""Times the ResBlock-tail / fork mixes (tcgen05 channel mix with an epilogue addend) against their HBM bound.
    python tools/bench_addend.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from deepsphere_weather_b200 import functional as F_  # noqa: E402
from deepsphere_weather_b200 import layers as L  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    timed = bench._event_timer(dev)
    hbm = bench._peaks()[0]
    print(f"{'case':40s} {'us':>8s} {'GB/s':>8s} {'frac':>6s}")
    for B, V, Fin, Fout in [(32, 12288, 24, 128), (32, 12288, 256, 64), (32, 3072, 128, 256), (32, 3072, 512, 128)]:
        lin = L.NodeLinear(Fin, Fout).to(dev)
        x = torch.randn(B, V, Fin, device=dev)
        a = torch.randn(B, V, Fout, device=dev)
        g = torch.randn(B, V, Fout, device=dev)
        d = torch.randn(B, V, Fin, device=dev)
        rz = torch.ones(1, device=dev)
        N = B * V
        with torch.no_grad():
            t = timed(lambda: F_.linear_rezero(x, lin.weight, lin.bias, a, rz))
            by = 4 * N * (Fin + 2 * Fout)
            print(f"tail  {Fin:4d}->{Fout:4d} rows {N:7d}            {t * 1e6:8.1f} {by / t / 1e9:8.0f} {by / t / 1e9 / hbm:6.2f}")
            t = timed(lambda: F_.NodeLinearFunction.apply(x, lin.weight, lin.bias))
            by = 4 * N * (Fin + Fout)
            print(f"plain {Fin:4d}->{Fout:4d} rows {N:7d}            {t * 1e6:8.1f} {by / t / 1e9:8.0f} {by / t / 1e9 / hbm:6.2f}")
            st = F_._ForkState()

            class Ctx:
                state = st

            def fork_bwd():
                st.g, st.w = g, lin.weight
                return F_.ForkFunction.backward(Ctx, d)

            t = timed(fork_bwd)
            by = 4 * N * (Fout + 2 * Fin)
            print(f"fork-bwd {Fout:4d}->{Fin:4d} rows {N:7d} (+addend)  {t * 1e6:8.1f} {by / t / 1e9:8.0f} {by / t / 1e9 / hbm:6.2f}")
            t = timed(lambda: torch.add(d, d))
            by = 12 * N * Fin
            print(f"torch add {Fin:4d} ch rows {N:7d}             {t * 1e6:8.1f} {by / t / 1e9:8.0f} {by / t / 1e9 / hbm:6.2f}")


if __name__ == "__main__":
    main()
