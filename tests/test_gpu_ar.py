"""GPU tests of the autoregressive training step (deepsphere_weather_b200/ar.py, csrc/dsw_ar.cu)
against the plain-torch composition it replaces (history shift by ``cat``, expand of the static fields, three-way ``cat``
along the feature axis; xforecasting's loop restated — the package itself is not installable)."""
import pytest
import torch

from _util import REL_TOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(lib):
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (there is no CPU fallback)")
    return torch.device("cuda:0")


def _torch_stack(dyn, bc, static):
    B, V = dyn[0].shape[:2]
    parts = [torch.stack(list(dyn), 1)]
    if bc:
        parts.append(torch.stack(list(bc), 1))
    if static is not None:
        parts.append(static[None, None].expand(B, len(dyn), V, static.shape[1]))
    return torch.cat(parts, dim=3)


@pytest.mark.parametrize("B,T,V,Fd,Fb,Fs", [(3, 3, 48, 2, 1, 4), (2, 1, 130, 5, 0, 0), (1, 4, 7, 1, 3, 0), (2, 2, 33, 2, 0, 3)])
def test_ar_stack_matches_torch_cat_bit_exact(B, T, V, Fd, Fb, Fs, dev):
    from deepsphere_weather_b200.ar import ar_stack

    torch.manual_seed(B * 10 + T)
    big = torch.randn(B, T + 1, V, Fd, device=dev)                       # slots are strided views of a history tensor
    dyn = [big[:, t].clone().requires_grad_(t % 2 == 0) if t < 2 else big[:, t] for t in range(T)]
    bc = [torch.randn(B, V, Fb, device=dev) for _ in range(T)] if Fb else None
    static = torch.randn(V, Fs, device=dev) if Fs else None
    X = ar_stack(dyn, bc, static)
    want = _torch_stack([d.detach() for d in dyn], bc, static)
    assert X.shape == (B, T, V, Fd + Fb + Fs) and torch.equal(X, want)
    g = torch.randn_like(X)
    X.backward(g)
    for t, d in enumerate(dyn):
        if d.requires_grad:
            assert torch.equal(d.grad, g[:, t, :, :Fd])


def _rollout_reference(model, crit, history, bc, static, targets, weights):
    T = history.shape[1]
    hist = history
    total = 0.0
    for i, w in enumerate(weights):
        X = _torch_stack([hist[:, t] for t in range(T)], [bc[:, i + t] for t in range(T)] if bc is not None else None, static)
        pred = model(X)[:, 0]
        total = total + w * crit(pred, targets[:, i])
        hist = torch.cat((hist[:, 1:], pred.unsqueeze(1)), dim=1)        # the shifted history, materialised
    return total


def test_ar_rollout_matches_torch_composition_and_replays_from_a_cuda_graph(dev):
    from deepsphere_weather_b200 import models as M
    from deepsphere_weather_b200.ar import ARRollout
    from deepsphere_weather_b200.ddp import FlatGradBucket
    from deepsphere_weather_b200.losses import WeightedMSELoss

    nside, B, T, Fd, Fb, Fs, ar_it = 8, 2, 3, 2, 1, 4, 2
    V = 12 * nside * nside
    model = M.UNetSpherical(M.default_tensor_info(V, input_n_feature=Fd + Fb + Fs, output_n_feature=Fd, input_n_time=T), "healpix",
                            {"subdivisions": nside, "nest": True}, kernel_size_conv=3, pool_method="interp")
    M.deterministic_fill(model, seed=2, rezero=1.0)
    model = model.to(dev)
    bucket = FlatGradBucket(model)
    torch.manual_seed(9)
    crit = WeightedMSELoss(weights=(torch.rand(V, device=dev) + 0.5))
    weights = [0.5, 0.3, 0.2]
    roll = ARRollout(model, crit, ar_it, weights)

    def batch(seed):
        g = torch.Generator(device="cpu").manual_seed(seed)
        mk = lambda *s: torch.randn(*s, generator=g).to(dev)
        return mk(B, T, V, Fd), mk(B, T + ar_it + 1, V, Fb), mk(V, Fs), mk(B, ar_it + 1, V, Fd)

    args = batch(1)
    loss = roll.step(*args, zero_grad=bucket.zero_)
    got = bucket.flat.clone()
    bucket.zero_()
    ref = _rollout_reference(model, crit, *args, weights)
    ref.backward()
    ref_val = ref.item()
    # (a loss tensor kept alive keeps its AccumulateGrad nodes — bound to the default stream — alive too, and a later
    # capture would route the parameter gradients through them: PyTorch's CUDA-graph rule, not this library's)
    del ref
    assert abs(loss.item() - ref_val) <= 1e-6 * abs(ref_val)
    assert rel_err(got, bucket.flat) < REL_TOL          # same kernels, different stacking: fp32 accumulation order only

    roll.capture(*args, zero_grad=bucket.zero_)
    for seed in (2, 3):
        new = batch(seed)
        l_graph = roll.replay(*new).clone()
        g_graph = bucket.flat.clone()
        l_eager = roll.step(*new, zero_grad=bucket.zero_)
        assert torch.equal(l_graph, l_eager)
        assert torch.equal(g_graph, bucket.flat)         # the captured rollout is the same launches: bit-identical
