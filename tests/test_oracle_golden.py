"""CPU: the oracle (oracle/cheb_oracle.py) against the golden vectors produced by the unmodified
reference (oracle/make_golden.py), and — when /root/reference is present — against the reference
run live.  This is what pins the oracle."""
import numpy as np
import pytest
import torch
from scipy import sparse

from _util import CONV_CASES, UNET_CASES, coo_from, golden, rel_err
from deepsphere_weather_b200 import graphs as G
from deepsphere_weather_b200 import models as M
from oracle import cheb_oracle as O
from oracle import ref_import
from oracle import unet_oracle as U


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_oracle_matches_reference_golden(case):
    g = golden(case)
    lap = coo_from(g, "lap")
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    w = torch.from_numpy(g["w"]).requires_grad_(True)
    b = torch.from_numpy(g["b"]).requires_grad_(True) if "b" in g.files else None
    y = O.conv_cheb_layer(lap, x, w, b)
    y.backward(torch.from_numpy(g["dy"]))
    assert rel_err(y, g["y"]) < 2e-6
    assert rel_err(x.grad, g["dx"]) < 2e-6
    assert rel_err(w.grad, g["dw"]) < 2e-6
    if b is not None:
        assert rel_err(b.grad, g["db"]) < 2e-6


@pytest.mark.parametrize("case", ["conv_cfg1", "conv_k1", "conv_k2_nobias", "conv_last_layer"])
def test_dense_f64_restatement_agrees(case):
    g = golden(case)
    lap = coo_from(g, "lap").to_dense().numpy()
    y = O.conv_cheb_dense_f64(lap, g["x"], g["w"], g["b"] if "b" in g.files else None)
    assert rel_err(y, g["y"]) < 5e-6


def test_known_answers():
    # K = 1: y = x @ W[:,0,:] + b, no Laplacian involved (layers.py:163,208)
    g = golden("conv_k1")
    y = g["x"] @ g["w"][:, 0, :] + g["b"]
    assert rel_err(y, g["y"]) < 2e-6
    # L = c*I  =>  T_k = cos(k*acos(c)) * I
    n, c = 12, 0.3
    lap = torch.sparse_coo_tensor(torch.arange(n).repeat(2, 1), torch.full((n,), c), (n, n)).coalesce()
    x = torch.randn(2, n, 3)
    w = torch.randn(3, 5, 4)
    coef = torch.tensor([np.cos(k * np.arccos(c)) for k in range(5)], dtype=torch.float32)
    expect = torch.einsum("bvf,k,fko->bvo", x, coef, w)
    assert rel_err(O.conv_cheb(lap, x, w), expect) < 1e-5


def test_shape_error_matches_reference():
    lap = G.healpix_laplacian(1)
    with pytest.raises(ValueError, match="Input tensor shape does not match"):
        O.conv_cheb(lap, torch.zeros(1, 12, 3), torch.zeros(4, 2, 5))


def _pool_mats(g):
    pool = sparse.coo_matrix((g["pool_dat"], (g["pool_row"], g["pool_col"])), shape=(48, 192))
    unpool = sparse.coo_matrix((g["unpool_dat"], (g["unpool_row"], g["unpool_col"])), shape=(192, 48))
    return pool, unpool


def test_pool_oracles_match_reference_golden():
    g = golden("pools")
    x = torch.from_numpy(g["x"])
    pool_m, unpool_m = _pool_mats(g)
    # interpolation remap
    P, Uq = coo_from(g, "interp_pool"), coo_from(g, "interp_unpool")
    yp = O.remap(P, x)
    assert rel_err(yp, g["interp_pooled"]) < 2e-6
    assert rel_err(O.remap(Uq, yp), g["interp_unpooled"]) < 2e-6
    # max-area one-hot matrices: bit-exact structure
    Pa = O.max_area_pool_matrix(sparse.csr_matrix(pool_m))
    Ua = O.max_area_unpool_matrix(sparse.csr_matrix(pool_m.T))
    assert torch.equal(Pa.indices(), coo_from(g, "maxarea_pool").indices())
    assert torch.equal(Ua.indices(), coo_from(g, "maxarea_unpool").indices())
    assert np.array_equal(O.remap(Pa, x).numpy(), g["maxarea_pooled"])
    # max-value pool: values and int64 indices bit-exact
    M_ = torch.sparse_coo_tensor(torch.from_numpy(np.stack([pool_m.row, pool_m.col]).astype(np.int64)),
                                 torch.from_numpy(pool_m.data.astype(np.float32)), pool_m.shape).coalesce()
    yp, idx = O.maxval_pool(M_, x)
    assert np.array_equal(yp.numpy(), g["maxval_pooled"])
    assert np.array_equal(idx.numpy(), g["maxval_index"])
    assert np.array_equal(O.maxval_unpool(192, yp, idx).numpy(), g["maxval_unpooled"])
    # nested pools
    yp, idx = O.healpix_max_pool(x, 4)
    assert np.array_equal(yp.numpy(), g["hmax_pooled"]) and np.array_equal(idx.numpy(), g["hmax_index"])
    assert np.array_equal(O.healpix_max_unpool(yp, idx, 4).numpy(), g["hmax_unpooled"])
    ya, none = O.healpix_avg_pool(x, 4)
    assert none is None and np.array_equal(ya.numpy(), g["havg_pooled"])
    assert np.array_equal(O.healpix_avg_unpool(ya, 4).numpy(), g["havg_unpooled"])


def test_nested_pool_invariants():
    # tutorials/interpolation_pooling.ipynb cell 16; layers.py:562
    pool, unpool = G.nested_pool_matrices(192, 4)
    assert np.allclose(np.asarray(pool.sum(1)).ravel(), 1.0)
    assert np.allclose(np.asarray(unpool.sum(1)).ravel(), 1.0)
    assert np.allclose((pool @ unpool).toarray(), np.eye(48))
    pool, unpool = G.random_overlap_pool_matrices(192, 48, seed=3)
    assert np.allclose(np.asarray(pool.sum(1)).ravel(), 1.0)
    assert np.allclose(np.asarray(unpool.sum(1)).ravel(), 1.0)


@pytest.mark.parametrize("name,pool_method,K,seed", UNET_CASES)
def test_unet_oracle_matches_reference_golden(name, pool_method, K, seed):
    g = golden(name)
    laps = [coo_from(g, f"lap{i}") for i in range(3)]
    model = U.build_unet_oracle(M.default_tensor_info(768), "healpix", {"subdivisions": 8, "nest": True},
                                kernel_size_conv=K, pool_method=pool_method, laplacians=laps)
    U.fill_parameters(model, seed)
    y = model(torch.from_numpy(g["x"]))
    assert rel_err(y, g["y"]) < 1e-5
    loss = (y**2).mean()
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    grads = dict(model.named_parameters())
    names = [str(n) for n in g["grad_names"]]
    assert names == sorted(n for n, _ in model.named_parameters())
    for n, ref_norm in zip(names, g["grad_norms"]):
        got = grads[n].grad.norm().item()
        assert abs(got - ref_norm) <= 1e-4 * max(ref_norm, 1e-6) + 1e-9, n
    for key in g.files:
        if key.startswith("grad__"):
            assert rel_err(grads[key[6:]].grad, g[key]) < 1e-4, key


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present (GPU box)")
def test_oracle_against_live_reference():
    ref_layers, _ = ref_import.load_reference()
    torch.manual_seed(3)
    lap = G.healpix_laplacian(4)
    x = torch.randn(2, 192, 7)
    w = torch.randn(7, 5, 9) * 0.3
    assert torch.equal(ref_layers.conv_cheb(lap, x, w), O.conv_cheb(lap, x, w))
    pool_m, _ = G.random_overlap_pool_matrices(192, 48, seed=9)
    ref_pool = ref_layers.GeneralMaxValPool(pool_m)
    yr, ir = ref_pool(x)
    yo, io = O.maxval_pool(ref_pool.remap_matrix, x)
    assert torch.equal(yr, yo) and torch.equal(ir, io)


# ----------------------------------------------------------------------------------------------
# Training-loop loss: oracle restatement against the unmodified reference
# ----------------------------------------------------------------------------------------------


def test_weighted_mse_oracle_matches_reference_golden():
    from oracle import loss_oracle as LO

    g = golden("wmse")
    for tag in ("a", "b"):
        w = torch.from_numpy(g[f"{tag}_w"])
        label = torch.from_numpy(g[f"{tag}_label"])
        for red in ("mean", "sum", "none"):
            for use_w in (True, False):
                pred = torch.from_numpy(g[f"{tag}_pred"]).requires_grad_(True)
                key = f"{tag}_{red}_{'w' if use_w else 'u'}"
                val = LO.weighted_mse(pred, label, w if use_w else None, red)
                assert rel_err(val, g[key]) < 1e-6, key
                if red != "none":
                    val.backward()
                    assert rel_err(pred.grad, g[key + "_grad"]) < 1e-6, key
    y = torch.from_numpy(g["reshape_in"])
    assert torch.equal(LO.reshape_4_loss(y, ["sample", "time", "node", "feature"]), torch.from_numpy(g["reshape_out"]))
    with pytest.raises(ValueError):
        LO.weighted_mse(torch.zeros(1, 4, 1), torch.zeros(1, 4, 1), torch.ones(3))


def test_reshape_tensors_4_loss_matches_reference_golden():
    from deepsphere_weather_b200.losses import reshape_tensors_4_loss

    g = golden("wmse")
    y = torch.from_numpy(g["reshape_in"])
    a, b = reshape_tensors_4_loss(y, y + 1, {"sample": 0, "time": 1, "node": 2, "feature": 3})
    assert torch.equal(a, torch.from_numpy(g["reshape_out"])) and torch.equal(b, a + 1)


# ----------------------------------------------------------------------------------------------
# Equiangular image path: the thin torch mirrors against the unmodified reference
# ----------------------------------------------------------------------------------------------


@pytest.mark.skipif(not ref_import.reference_available(), reason="the unmodified reference is only present in the build container")
@pytest.mark.parametrize("periodic", [True, False])
def test_equiangular_image_layers_match_the_reference(periodic):
    import torch

    from deepsphere_weather_b200 import layers_equiangular as E

    ref_layers, _ = ref_import.load_reference()
    torch.manual_seed(3)
    n_lat, n_lon, B, Fin, Fout = 8, 16, 2, 5, 7
    x = torch.randn(B, n_lat * n_lon, Fin)
    ref = ref_layers.Conv2dEquiangular(Fin, Fout, 3, lonlat_ratio=2, periodic_padding=periodic, bias=True)
    ours = E.Conv2dEquiangular(Fin, Fout, 3, lonlat_ratio=2, periodic_padding=periodic, bias=True)
    assert list(ref.state_dict()) == list(ours.state_dict())
    ours.load_state_dict(ref.state_dict(), strict=True)
    assert torch.allclose(ours(x), ref(x), atol=1e-6)
    for tag in ("max", "avg"):
        rp, ru = (cls(lonlat_ratio=2, kernel_size=4) for cls in ref_layers.ALL_POOL["equiangular"][tag])
        op, ou = (cls(lonlat_ratio=2, kernel_size=4) for cls in E.EQUIANGULAR_POOL[tag])
        (yr, ir), (yo, io) = rp(x), op(x)
        assert torch.equal(yo, yr) and (ir is None) == (io is None) and (ir is None or torch.equal(io, ir))
        assert torch.equal(ou(yo, io), ru(yr, ir))


@pytest.mark.skipif(not ref_import.reference_available(), reason="the unmodified reference is only present in the build container")
@pytest.mark.parametrize("pool_method", ["max", "avg"])
def test_image_unet_matches_the_reference_on_the_equiangular_grid(pool_method):
    """conv_type="image" + equiangular index pools: the whole reference architecture on the dense lat x lon path."""
    import torch

    _, ref_models = ref_import.load_reference()
    nlat, nlon, B = 16, 32, 2
    V = nlat * nlon
    kw = dict(kernel_size_conv=3, conv_type="image", pool_method=pool_method, periodic_padding=True)
    args = (M.default_tensor_info(V), "equiangular", {"nlat": nlat, "nlon": nlon})
    ref = ref_models.UNetSpherical(*args, **kw)
    # the image path is plain PyTorch and runs on the CPU too; the ResBlock tail then has to be the reference's own
    # (torch.nn.Linear + in-place ops) instead of the CUDA-only fused kernels of the default backend
    from types import SimpleNamespace

    from deepsphere_weather_b200 import layers as L

    backend = SimpleNamespace(ConvCheb=L.ConvCheb, healpix_pools={}, general_pools=L.PoolUnpoolBlock.getGeneralPoolUnpoolLayer)
    ours = M.UNetSpherical(*args, backend=backend, **kw)
    sd = {k: v for k, v in ref.state_dict().items() if not v.is_sparse}
    missing = ours.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all("laplacian" in k for k in missing.missing_keys)
    torch.manual_seed(0)
    with torch.no_grad():
        for m in (ref, ours):
            for n, p in m.named_parameters():
                if n.endswith("rezero_weight"):
                    p.fill_(0.7)
    x = torch.randn(B, 3, V, 7)
    yr, yo = ref(x), ours(x)
    assert torch.allclose(yo, yr, rtol=1e-5, atol=1e-6)
    yr.square().mean().backward()
    yo.square().mean().backward()
    for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
        assert torch.allclose(p.grad, q.grad, rtol=1e-4, atol=1e-7), n
