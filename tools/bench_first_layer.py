"""Timing of the two boundary layers of the U-Net (21 -> 64 and 64 -> 2 at nside 32, B 32, K 4) under both evaluation
orders, and of the SpMM recurrence at 24 / 32 / 64 channels (run under gpurun)."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepsphere_weather_b200 import _lib, functional as F_, graphs as G, layers as L

def timed(fn, flush, iters=8, warm=3):
    ts = []
    for i in range(warm + iters):
        flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        if i >= warm: ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)

dev = torch.device("cuda:0"); lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lap = G.healpix_laplacian(32); plan = F_.plan_for(lap.to(dev)); V = lap.shape[0]; B = 32
for F in (24, 32, 64):
    x = torch.randn(B, V, F, device=dev)
    for nochain in (0, 1):
        lib.dsw_set_option(17, nochain)
        print(f"terms F={F:3d} K=4 {'hop-by-hop' if nochain else 'auto      '} {timed(lambda: F_.cheb_terms(x, plan, 4), flush):8.1f} us", flush=True)
lib.dsw_set_option(17, 0)
for Fin, Fout in ((21, 64), (64, 2)):
    layer = L.ConvCheb(Fin, Fout, 4, lap).to(dev)
    x = torch.randn(B, V, Fin, device=dev)
    xg = x.clone().requires_grad_(Fin != 21)
    dy = torch.randn(B, V, Fout, device=dev)
    for fa in (1, 2):
        for ba in (1, 2):
            lib.dsw_set_option(4, fa); lib.dsw_set_option(5, ba)
            with torch.no_grad():
                tf = timed(lambda: layer(x), flush)
            def fb():
                layer.zero_grad(set_to_none=True); xg.grad = None
                layer(xg).backward(dy)
            tb = timed(fb, flush)
            print(f"layer {Fin}->{Fout} fwd_algo {fa} bwd_algo {ba}: fwd {tf:8.1f} us  fwd+bwd {tb:8.1f} us", flush=True)
    lib.dsw_set_option(4, 0); lib.dsw_set_option(5, 0)
    with torch.no_grad():
        tf = timed(lambda: layer(x), flush)
    print(f"layer {Fin}->{Fout} auto: fwd {tf:8.1f} us  fwd+bwd {timed(fb, flush):8.1f} us  (algos {lib.dsw_cheb_fwd_algo(max(Fin,24) if Fin==21 else Fin, max(Fout,4), 4)}, {lib.dsw_cheb_bwd_algo(max(Fin,24) if Fin==21 else Fin, max(Fout,4), 4)})", flush=True)
