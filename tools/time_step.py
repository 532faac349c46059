"""Device time of the cfg3 U-Net step (forward + MSE + zero-grad + backward), for same-box A/B runs with DSW_OPTIONS / DSW_LIB_PATH.

    python tools/time_step.py [steps] [nside] [batch]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    nside = int(sys.argv[2]) if len(sys.argv) > 2 else bench.NSIDE
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else bench.BATCH_PER_GPU
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model, V = bench.build_model(dev, nside=nside)
    x = torch.randn(batch, 3, V, 7, device=dev)
    y = torch.randn(batch, 1, V, 2, device=dev)
    crit = torch.nn.MSELoss()

    def step():
        loss = crit(model(x), y)
        model.zero_grad(set_to_none=True)
        loss.backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"{os.environ.get('DSW_LIB_PATH', 'libdsw.so')} DSW_OPTIONS={os.environ.get('DSW_OPTIONS', '')}: "
          f"nside {nside} B {batch}: {e0.elapsed_time(e1) / steps:.3f} ms / step")


if __name__ == "__main__":
    main()
