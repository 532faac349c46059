"""cfg5 (BASELINE.json configs[4]): equiangular 400 x 200 sampling (80 000 nodes), k-NN-20 Laplacian with
highly irregular degree near the poles, ConvCheb K = 6, Cin = Cout = 128: time of the SpMM recurrence and of
the layer forward / forward + backward.

    python tools/bench_cfg5.py [B]
"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepsphere_weather_b200 import functional as F_  # noqa: E402
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
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device("cuda:0")
    F, K = 128, 6
    lap = G.prepare_torch_laplacian(G.knn_laplacian(G.equiangular_xyz(200, 400), 20))
    V = lap.shape[0]
    c = lap.coalesce()
    deg = torch.bincount(c.indices()[0], minlength=V)
    print(f"V {V} nnz {c.values().numel()} degree min {int(deg.min())} mean {float(deg.float().mean()):.1f} max {int(deg.max())}")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    plan = F_.plan_for(lap.to(dev))
    x = torch.randn(B, V, F, device=dev)
    t = timed(lambda: F_.cheb_terms(x, plan, K), flush)
    alg = 4 * B * V * F * K + plan.operand_bytes
    print(f"terms (K-1 = {K - 1} hops) B {B} F {F}: {t:8.1f} us  {alg / t / 1e3:7.1f} GB/s algorithmic")
    layer = L.ConvCheb(F, F, K, lap).to(dev)
    with torch.no_grad():
        tf = timed(lambda: layer(x), flush)
    xg = x.clone().requires_grad_(True)
    dy = torch.randn(B, V, F, device=dev)

    def fb():
        layer.zero_grad(set_to_none=True)
        xg.grad = None
        layer(xg).backward(dy)

    tb = timed(fb, flush)
    print(f"ConvCheb 128->128 K6 fwd {tf:8.1f} us  fwd+bwd {tb:8.1f} us  ({B * V * F / tf * 1e6 / 1e9:.1f} G node-channels/s fwd)")


if __name__ == "__main__":
    main()
