"""Diagnose the tcgen05 weight-gradient kernel: compare against torch on small shapes and print the
structure of the mismatch (row / column permutations, scale)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepsphere_weather_b200 import _lib, functional as F_, graphs as G

lib = _lib.load()
dev = torch.device("cuda:0")

def run(B, nside, Fin, Fout, K, mode):
    lib.dsw_set_mix_mode(mode)
    torch.manual_seed(0)
    lap = G.healpix_laplacian(nside).to(dev)
    plan = F_.plan_for(lap)
    V = lap.shape[0]
    x = torch.randn(B, V, Fin, device=dev)
    dy = torch.randn(B, V, Fout, device=dev)
    dw = torch.full((Fin, K, Fout), float("nan"), device=dev)
    db = torch.full((Fout,), float("nan"), device=dev)
    ws = torch.empty(lib.dsw_cheb_bwd_weight_workspace_bytes(B, V, Fin, Fout, K), dtype=torch.uint8, device=dev)
    rc = lib.dsw_cheb_bwd_weight(plan.handle, x.data_ptr(), x.stride(0), x.stride(1), dy.data_ptr(), dw.data_ptr(),
                                 db.data_ptr(), B, Fin, Fout, K, ws.data_ptr(), ws.numel(),
                                 torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == 0, (rc, lib.dsw_strerror(rc), lib.dsw_last_cuda_error_string())
    return x, dy, dw.cpu(), db.cpu(), plan

def analyse(name, got, ref):
    got, ref = got.double().numpy(), ref.double().numpy()
    err = np.abs(got - ref).max() / np.abs(ref).max()
    print(f"  {name}: rel err {err:.3e}  nan={np.isnan(got).sum()}  |got|max={np.nanmax(np.abs(got)):.3f} |ref|max={np.abs(ref).max():.3f}")
    return err

for (B, nside, Fin, Fout, K) in [(4, 2, 64, 64, 1), (4, 2, 128, 64, 1), (4, 2, 64, 128, 1), (4, 2, 64, 64, 2), (3, 2, 8, 12, 1)]:
    print(f"case B={B} V={12*nside*nside} Fin={Fin} Fout={Fout} K={K}")
    x, dy, dw0, db0, plan = run(B, nside, Fin, Fout, K, 0)
    _, _, dw1, db1, _ = run(B, nside, Fin, Fout, K, 1)
    analyse("simt dW vs tc dW", dw1, dw0)
    analyse("simt db vs tc db", db1, db0)
    if K == 1:
        ref = torch.einsum("bvf,bvo->fo", x.cpu().double(), dy.cpu().double())
        g = dw1[:, 0, :].double()
        e = analyse("tc dW vs einsum", g.float(), ref.float())
        if e > 1e-3:
            # look for row / column permutations
            gn = g / (g.norm(dim=1, keepdim=True) + 1e-30); rn = ref / ref.norm(dim=1, keepdim=True)
            sim = gn @ rn.T
            best = sim.abs().argmax(1)
            print("   row map (tc row i ~ ref row):", best[:16].tolist(), " sim:", [round(float(sim[i, best[i]]), 3) for i in range(8)])
            gc = g.T / (g.T.norm(dim=1, keepdim=True) + 1e-30); rc_ = ref.T / ref.T.norm(dim=1, keepdim=True)
            simc = gc @ rc_.T
            bestc = simc.abs().argmax(1)
            print("   col map:", bestc[:16].tolist(), " sim:", [round(float(simc[i, bestc[i]]), 3) for i in range(8)])
            print("   tc[:3,:6] ", g[:3, :6].numpy().round(3).tolist())
            print("   ref[:3,:6]", ref[:3, :6].numpy().round(3).tolist())
            # partial sums hypothesis: only some rows n contribute?
            N = x.shape[0] * x.shape[1]
            xf, df = x.cpu().double().reshape(N, -1), dy.cpu().double().reshape(N, -1)
            for lo, hi in [(0, 8), (0, 16), (0, 64), (0, 128), (64, 128)]:
                part = xf[lo:hi].T @ df[lo:hi]
                print(f"   rows[{lo}:{hi}] only -> err {float((g-part).abs().max()/ref.abs().max()):.3e}")
