"""Driver for ncu captures of the short-K channel mixes (ResBlock tail with addend, plain Linear, fork backward).
    ncu ... python tools/run_tail_once.py [Fin Fout [rows_V]]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepsphere_weather_b200 import functional as F_  # noqa: E402
from deepsphere_weather_b200 import layers as L  # noqa: E402


def main():
    Fin = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    Fout = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    V = int(sys.argv[3]) if len(sys.argv) > 3 else 12288
    dev = torch.device("cuda:0")
    B = 32
    lin = L.NodeLinear(Fin, Fout).to(dev)
    x = torch.randn(B, V, Fin, device=dev)
    a = torch.randn(B, V, Fout, device=dev)
    rz = torch.ones(1, device=dev)
    with torch.no_grad():
        for _ in range(3):
            F_.NodeLinearFunction.apply(x, lin.weight, lin.bias)   # mix_tma_kernel<false>
            F_.linear_rezero(x, lin.weight, lin.bias, a, rz)       # mix_tma_kernel<true>
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
