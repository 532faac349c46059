"""In-situ kernel times of the cfg3 U-Net step from the CUPTI activity trace (torch.profiler): unlike the ncu launch list
(cold caches, serialised) these are the durations inside the running step, plus the GPU idle share of the step.

    python tools/step_kineto.py [steps] [nside] [batch]
"""
import os
import re
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
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
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    wall = e0.elapsed_time(e1) / 10
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
    tot = defaultdict(float)
    cnt = defaultdict(int)
    spans = []
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and ev.device_time > 0 and "Memcpy" not in ev.name and "Memset" not in ev.name:
            name = re.sub(r"\(.*", "", ev.name)
            name = re.sub(r"^void ", "", name)[:70]
            tot[name] += ev.device_time
            cnt[name] += 1
            spans.append((ev.time_range.start, ev.time_range.end))
    sub = os.environ.get("KINETO_LIST")
    for one in (sub.split(",") if sub else []):  # per-launch durations (us) of the kernels whose name contains it, first step
        evs = sorted((ev.time_range.start, ev.device_time, ev.name) for ev in prof.events()
                     if ev.device_type == torch.autograd.DeviceType.CUDA and one in ev.name)
        evs = evs[: len(evs) // steps]
        print("# per-launch", one, [round(d, 1) for _, d, _ in evs])
    spans.sort()
    busy, cur_s, cur_e = 0.0, None, None
    for s, e in spans:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        busy += cur_e - cur_s
    total = sum(tot.values())
    print(f"# eager step (no profiler): {wall:.3f} ms;  kernel time sum {total / steps / 1e3:.3f} ms / step;  GPU busy (union) "
          f"{busy / steps / 1e3:.3f} ms / step  over {steps} steps")
    for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{t / steps:10.1f} us {100 * t / total:5.1f}%  n={cnt[name] / steps:6.1f}  {name}")


if __name__ == "__main__":
    main()
