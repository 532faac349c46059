"""TEST INFRASTRUCTURE ONLY — generate ``tests/golden/*.npz`` by running the UNMODIFIED reference.

Run in the build container (needs ``/root/reference``):

    python -m oracle.make_golden

The reference has no tests or golden vectors, so these files *are* the pin: inputs
and outputs of the reference's own ``conv_cheb`` / ``ConvCheb`` / pool modules /
``UNetSpherical`` on seeded inputs.  They travel to the GPU box, where ``/root/reference`` does
not exist.  Sparse operators are stored explicitly (their construction is non-deterministic in
the reference, SURVEY.md §0.4, and is an input to the path anyway).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deepsphere_weather_b200 import graphs as G  # noqa: E402
from deepsphere_weather_b200 import models as M  # noqa: E402
from oracle import ref_import  # noqa: E402
from oracle.unet_oracle import fill_parameters  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _coo_arrays(t: torch.Tensor, prefix: str) -> dict:
    t = t.coalesce()
    return {
        f"{prefix}_idx": t.indices().numpy().astype(np.int32),
        f"{prefix}_val": t.values().numpy().astype(np.float32),
        f"{prefix}_shape": np.asarray(t.shape, dtype=np.int64),
    }


def _rng(seed):
    return np.random.default_rng(seed)


def _t(a, grad=False):
    return torch.from_numpy(np.ascontiguousarray(a.astype(np.float32))).requires_grad_(grad)


def conv_case(ref_layers, name, nside, B, Fin, Fout, K, seed, bias=True):
    r = _rng(seed)
    V = 12 * nside * nside
    L = G.healpix_laplacian(nside)
    x = _t(r.standard_normal((B, V, Fin)), True)
    layer = ref_layers.ConvCheb(Fin, Fout, K, L, bias=bias)
    with torch.no_grad():
        layer.weight.copy_(_t(r.standard_normal((Fin, K, Fout)) * np.sqrt(2.0 / (Fin * K))))
        if bias:
            layer.bias.copy_(_t(r.standard_normal(Fout) * 0.1))
    dy = _t(r.standard_normal((B, V, Fout)))
    y = layer(x)
    y.backward(dy)
    out = dict(x=x.detach().numpy(), w=layer.weight.detach().numpy(), dy=dy.numpy(), y=y.detach().numpy(),
               dx=x.grad.numpy(), dw=layer.weight.grad.numpy(), **_coo_arrays(L, "lap"))
    if bias:
        out.update(b=layer.bias.detach().numpy(), db=layer.bias.grad.numpy())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "y absmax", float(np.abs(out["y"]).max()))


def pool_cases(ref_layers):
    r = _rng(100)
    B, V, Vc, F = 3, 192, 48, 10
    x_np = r.standard_normal((B, V, F)).astype(np.float32)
    x_np[0, :8, 0] = 1.25  # ties inside pooling windows: first maximum must win
    out = dict(x=x_np)

    # --- random overlapping ("Voronoi-like") interpolation matrices
    pool_m, unpool_m = G.random_overlap_pool_matrices(V, Vc, seed=3)
    for tag, cls_pool, cls_unpool, unpool_arg in (
        ("interp", ref_layers.GeneralAvgPool, ref_layers.GeneralAvgUnpool, unpool_m),
        ("maxarea", ref_layers.GeneralMaxAreaPool, ref_layers.GeneralMaxAreaUnpool, pool_m.T),
    ):
        pool, unpool = cls_pool(pool_m), cls_unpool(unpool_arg)
        x = _t(x_np, True)
        yp, _ = pool(x)
        yu = unpool(yp)
        g = _t(r.standard_normal(tuple(yu.shape)))
        yu.backward(g)
        out.update({f"{tag}_pooled": yp.detach().numpy(), f"{tag}_unpooled": yu.detach().numpy(),
                    f"{tag}_g": g.numpy(), f"{tag}_dx": x.grad.numpy(),
                    **_coo_arrays(pool.remap_matrix, f"{tag}_pool"), **_coo_arrays(unpool.remap_matrix, f"{tag}_unpool")})
    out.update(pool_row=pool_m.row.astype(np.int32), pool_col=pool_m.col.astype(np.int32),
               pool_dat=pool_m.data.astype(np.float64), unpool_row=unpool_m.row.astype(np.int32),
               unpool_col=unpool_m.col.astype(np.int32), unpool_dat=unpool_m.data.astype(np.float64))

    # --- max-value pooling with indices
    pool, unpool = ref_layers.GeneralMaxValPool(pool_m), ref_layers.GeneralMaxValUnpool(unpool_m)
    x = _t(x_np, True)
    yp, idx = pool(x)
    yu = unpool(yp, idx)
    g = _t(r.standard_normal(tuple(yu.shape)))
    yu.backward(g)
    out.update(maxval_pooled=yp.detach().numpy(), maxval_index=idx.numpy().astype(np.int64),
               maxval_unpooled=yu.detach().numpy(), maxval_g=g.numpy(), maxval_dx=x.grad.numpy())

    # --- nested HEALPix pools
    for tag, cls_pool, cls_unpool in (("hmax", ref_layers.HealpixMaxPool, ref_layers.HealpixMaxUnpool),
                                      ("havg", ref_layers.HealpixAvgPool, ref_layers.HealpixAvgUnpool)):
        pool, unpool = cls_pool(kernel_size=4), cls_unpool(kernel_size=4)
        x = _t(x_np, True)
        yp, idx = pool(x)
        yu = unpool(yp, idx)
        g = _t(r.standard_normal(tuple(yu.shape)))
        yu.backward(g)
        out.update({f"{tag}_pooled": yp.detach().numpy(), f"{tag}_unpooled": yu.detach().numpy(),
                    f"{tag}_g": g.numpy(), f"{tag}_dx": x.grad.numpy()})
        if idx is not None:
            out[f"{tag}_index"] = idx.numpy().astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "pools.npz"), **out)
    print("pools ok")


def unet_case(ref_layers, ref_models, name, pool_method, K, seed):
    nside, B = 8, 2
    V = 12 * nside * nside
    ti = M.default_tensor_info(V)
    if pool_method in ("interp", "maxval", "maxarea"):
        # xsphere/CDO is absent: hand the reference exact nested-pixel matrices through its own
        # build_pooling_matrices hook (monkey-patched at run time; reference files untouched).
        ref_layers.build_pooling_matrices = lambda src, dst: G.nested_pool_matrices(src.n_vertices, 4)
    model = ref_models.UNetSpherical(ti, "healpix", {"subdivisions": nside, "nest": True}, kernel_size_conv=K,
                                     pool_method=pool_method)
    fill_parameters(model, seed)
    r = _rng(seed + 1)
    x = _t(r.standard_normal((B, 3, V, 7)))
    y = model(x)
    loss = (y**2).mean()
    loss.backward()
    grads = {n: p.grad.numpy() for n, p in model.named_parameters()}
    keep = ["conv1.convblock1.conv.weight", "conv1.convblock1.conv.bias", "conv3.convblock2.conv.bias",
            "uconv1_final.convblock1.conv.weight", "uconv1_final.rezero_weight", "conv2.res_connection.weight",
            "uconv2.convblock1.conv.bias"]
    out = dict(x=x.numpy(), y=y.detach().numpy(), loss=np.float64(loss.item()),
               grad_names=np.array(sorted(grads)), grad_norms=np.array([np.linalg.norm(grads[n]) for n in sorted(grads)]))
    for n in keep:
        out["grad__" + n] = grads[n]
    for i, lap in enumerate(model.laplacians):
        out.update(_coo_arrays(lap, f"lap{i}"))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", loss.item())


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    ref_layers, ref_models = ref_import.load_reference()
    conv_case(ref_layers, "conv_cfg1", nside=8, B=4, Fin=16, Fout=16, K=3, seed=1)
    conv_case(ref_layers, "conv_first_layer", nside=4, B=3, Fin=21, Fout=64, K=4, seed=2)
    conv_case(ref_layers, "conv_last_layer", nside=4, B=3, Fin=64, Fout=2, K=4, seed=3)
    conv_case(ref_layers, "conv_k1", nside=2, B=2, Fin=8, Fout=12, K=1, seed=4)
    conv_case(ref_layers, "conv_k2_nobias", nside=2, B=2, Fin=5, Fout=7, K=2, seed=5, bias=False)
    conv_case(ref_layers, "conv_k6_wide", nside=4, B=1, Fin=128, Fout=128, K=6, seed=6)
    pool_cases(ref_layers)
    unet_case(ref_layers, ref_models, "unet_max_k3", "max", 3, seed=10)
    unet_case(ref_layers, ref_models, "unet_interp_k4", "interp", 4, seed=11)
    unet_case(ref_layers, ref_models, "unet_maxval_k3", "maxval", 3, seed=12)


if __name__ == "__main__":
    main()
