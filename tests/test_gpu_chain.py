"""GPU tests of the fused persistent multi-hop chain kernel (csrc/dsw_chain.cu).

The chain kernel evaluates the same recurrence (reference modules/layers.py:163-169) with the same per-row
arithmetic order as the hop-by-hop launches, so the two must agree BIT FOR BIT; both are also held to the
CPU restatement of the reference recurrence at the north-star tolerance (1e-4 relative).
"""
import pytest
import torch

from _util import REL_TOL, rel_err

pytestmark = pytest.mark.gpu

OPT_NO_CHAIN, OPT_CHAIN_L2, OPT_CHAIN_MIN_PASS = 17, 18, 19
HOP_BY_HOP, FUSED = 1, 2   # DSW_OPT_NO_CHAIN: 0 = automatic (fused from one tile per SM on), 1 = never, 2 = wherever supported


@pytest.fixture(scope="module")
def dev(lib):
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (there is no CPU fallback)")
    return torch.device("cuda:0")


def _terms_ref_cpu(lap, x, K):
    B, V, F = x.shape
    x0 = x.permute(1, 2, 0).reshape(V, F * B)
    t = [x0, torch.sparse.mm(lap, x0)]
    for _ in range(2, K):
        t.append(2 * torch.sparse.mm(lap, t[-1]) - t[-2])
    return [tk.reshape(V, F, B).permute(2, 0, 1) for tk in t[1:K]]


@pytest.mark.parametrize("nside,B,F,K", [
    (8, 5, 64, 4),      # 12 tiles, odd batch
    (8, 3, 96, 5),      # a partial second slab
    (4, 7, 128, 6),     # 3 tiles: every hop pass is smaller than the grid -> the dependency waits really block
    (16, 4, 64, 4),
    (8, 2, 512, 3),     # 8 slabs
    (2, 3, 16, 4),      # one partial tile (48 rows)
    (8, 4, 24, 4),      # the 24-channel first layer: half-width entry loop, checked accumulator init / stores
    (8, 2, 32, 3),      # a whole slab of exactly 32 channels
    (8, 4, 64, 9),      # 8 hops: two chain launches (DSW_CHAIN_MAX_HOPS = 7)
])
@pytest.mark.parametrize("group", ["default", "one-sample-groups", "ragged-groups"])
def test_chain_is_bit_identical_to_hop_by_hop(nside, B, F, K, group, dev, lib):
    from deepsphere_weather_b200 import functional as F_
    from deepsphere_weather_b200 import graphs as G

    torch.manual_seed(nside * 100 + B)
    lap = G.healpix_laplacian(nside)
    plan = F_.plan_for(lap.to(dev))
    V = lap.shape[0]
    x = torch.randn(B, V, F)
    xg = x.to(dev)
    try:
        lib.dsw_set_option(OPT_NO_CHAIN, HOP_BY_HOP)
        want = F_.cheb_terms(xg, plan, K)
        lib.dsw_set_option(OPT_NO_CHAIN, FUSED)
        if group == "one-sample-groups":
            lib.dsw_set_option(OPT_CHAIN_L2, 1)        # budget of one byte: S = 1 unless the pass floor raises it
            lib.dsw_set_option(OPT_CHAIN_MIN_PASS, 1)
        elif group == "ragged-groups":
            lib.dsw_set_option(OPT_CHAIN_L2, 2 * 3 * V * F * 4)  # two samples per group: B = 3, 5, 7 leave a ragged last group
            lib.dsw_set_option(OPT_CHAIN_MIN_PASS, 1)
        launches0 = lib.dsw_launch_count()
        got = F_.cheb_terms(xg, plan, K)
        n_launches = lib.dsw_launch_count() - launches0
        if nside >= 4:
            assert n_launches == (1 if K - 1 <= 7 else 2), "the recurrence must run as fused chain launches"
        assert torch.equal(got, want)
        ref = _terms_ref_cpu(lap, x, K)
        for k in range(K - 1):
            assert rel_err(got[k], ref[k]) < REL_TOL
    finally:
        for key in (OPT_NO_CHAIN, OPT_CHAIN_L2, OPT_CHAIN_MIN_PASS):
            lib.dsw_set_option(key, 0)


@pytest.mark.parametrize("fwd_algo,bwd_algo", [(1, 1), (1, 2), (2, 1), (2, 2)])
def test_chain_conv_fwd_bwd_bit_identical(fwd_algo, bwd_algo, dev, lib):
    """Both evaluation orders, both directions (the Clenshaw chain runs in place on its G planes, the adjoint
    chains use the transposed plan): fused chain == hop by hop, bit for bit."""
    from deepsphere_weather_b200 import graphs as G
    from deepsphere_weather_b200 import layers as L

    torch.manual_seed(11)
    lap = G.healpix_laplacian(8)
    B, V, Fin, Fout, K = 6, lap.shape[0], 96, 40, 5
    x, dy = torch.randn(B, V, Fin, device=dev), torch.randn(B, V, Fout, device=dev)
    layer = L.ConvCheb(Fin, Fout, K, lap).to(dev)
    res = []
    try:
        lib.dsw_set_option(4, fwd_algo)
        lib.dsw_set_option(5, bwd_algo)
        for no_chain in (HOP_BY_HOP, FUSED):
            lib.dsw_set_option(OPT_NO_CHAIN, no_chain)
            layer.zero_grad()
            xg = x.clone().requires_grad_(True)
            y = layer(xg, activation="relu")
            y.backward(dy)
            res.append((y.detach().clone(), xg.grad.clone(), layer.weight.grad.clone(), layer.bias.grad.clone()))
    finally:
        for key in (4, 5, OPT_NO_CHAIN):
            lib.dsw_set_option(key, 0)
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_chain_on_nonsymmetric_operator(dev, lib):
    """deps(t) comes from the operator's own structure: a non-symmetric operator (different halo in L and L^T)."""
    from deepsphere_weather_b200 import functional as F_
    from deepsphere_weather_b200 import graphs as G

    torch.manual_seed(3)
    lap = G.healpix_laplacian(8).coalesce()
    idx, val = lap.indices(), lap.values().clone()
    keep = (idx[0] <= idx[1]) | (torch.rand(val.numel()) < 0.5)   # drop half of the lower triangle
    val = val * (0.5 + torch.rand(val.numel()))
    A = torch.sparse_coo_tensor(idx[:, keep], 0.5 * val[keep], lap.shape).coalesce()
    plan = F_.plan_for(A.to(dev))
    B, V, F, K = 4, A.shape[0], 64, 5
    x = torch.randn(B, V, F)
    try:
        lib.dsw_set_option(OPT_NO_CHAIN, FUSED)
        got = F_.cheb_terms(x.to(dev), plan, K)
        lib.dsw_set_option(OPT_NO_CHAIN, HOP_BY_HOP)
        want = F_.cheb_terms(x.to(dev), plan, K)
    finally:
        lib.dsw_set_option(OPT_NO_CHAIN, 0)
    assert torch.equal(got, want)
    ref = _terms_ref_cpu(A, x, K)
    for k in range(K - 1):
        assert rel_err(got[k], ref[k]) < REL_TOL


def test_chain_replays_in_a_cuda_graph_and_survives_ring_wrap(dev, lib):
    """The chain kernel's claim counter and epoch live in device memory and are advanced by the last CTA to
    leave, so a captured launch (same sync set, same kernel arguments every replay) stays correct, and so do
    more back-to-back eager launches than there are sync sets."""
    from deepsphere_weather_b200 import functional as F_
    from deepsphere_weather_b200 import graphs as G

    torch.manual_seed(5)
    lap = G.healpix_laplacian(8).to(dev)
    plan = F_.plan_for(lap)
    B, V, F, K = 5, lap.shape[0], 64, 4
    x = torch.randn(B, V, F, device=dev)
    lib.dsw_set_option(OPT_NO_CHAIN, HOP_BY_HOP)
    want = F_.cheb_terms(x, plan, K)
    lib.dsw_set_option(OPT_NO_CHAIN, FUSED)
    for i in range(20):  # > DSW_CHAIN_SETS launches queued back to back
        got = F_.cheb_terms(x, plan, K)
    assert torch.equal(got, want)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            F_.cheb_terms(x, plan, K)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = F_.cheb_terms(x, plan, K)
    for rep in range(4):
        xn = torch.randn(B, V, F, device=dev)
        x.copy_(xn)
        g.replay()
        torch.cuda.synchronize()
        got = out.clone()
        lib.dsw_set_option(OPT_NO_CHAIN, HOP_BY_HOP)
        want = F_.cheb_terms(xn, plan, K)
        lib.dsw_set_option(OPT_NO_CHAIN, FUSED)
        assert torch.equal(got, want), f"replay {rep}"
    lib.dsw_set_option(OPT_NO_CHAIN, 0)


def test_chain_full_size_nside64_matches_hop_by_hop(dev, lib):
    """BASELINE config of the SpMM metric (nside 64, B 32, F 64, K 4): fused == hop by hop, bit for bit, and a
    slice of samples against the CPU recurrence."""
    from deepsphere_weather_b200 import functional as F_
    from deepsphere_weather_b200 import graphs as G

    torch.manual_seed(64)
    lap = G.healpix_laplacian(64)
    plan = F_.plan_for(lap.to(dev))
    B, V, F, K = 32, lap.shape[0], 64, 4
    x = torch.randn(B, V, F, device=dev)
    launches0 = lib.dsw_launch_count()
    got = F_.cheb_terms(x, plan, K)   # automatic mode: 768 tiles >= one per SM -> one fused launch
    assert lib.dsw_launch_count() - launches0 == 1
    lib.dsw_set_option(OPT_NO_CHAIN, HOP_BY_HOP)
    try:
        want = F_.cheb_terms(x, plan, K)
    finally:
        lib.dsw_set_option(OPT_NO_CHAIN, 0)
    assert torch.equal(got, want)
    ref = _terms_ref_cpu(lap, x[29:31].cpu(), K)
    for k in range(K - 1):
        assert rel_err(got[k, 29:31], ref[k]) < REL_TOL
