"""CPU-side checks: the C-ABI library loads and exports every symbol include/dsw.h declares (no
compute calls without a GPU), the product refuses to run without CUDA, and the batch-sharded
data-parallel host logic (flat gradient bucket + all-reduce) is correct at world_size 2 on gloo."""
import ctypes as C
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deepsphere_weather_b200 import _lib  # noqa: E402
from deepsphere_weather_b200 import build as dsw_build  # noqa: E402
from deepsphere_weather_b200.ddp import FlatGradBucket, shard_batch  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    dsw_build.build()
    return _lib.load()


def test_every_header_symbol_is_exported_and_bound(lib):
    declared = _lib.header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"libdsw.so does not export {name}"
        assert name in _lib.SIGNATURES, f"{name} is declared in dsw.h but has no ctypes signature"
    for name in _lib.SIGNATURES:
        assert name in declared, f"{name} is bound in _lib.py but not declared in include/dsw.h"


def test_version_and_error_strings(lib):
    assert lib.dsw_version() >= 100
    assert b"success" in lib.dsw_strerror(0)
    for code in range(-7, 0):
        assert len(lib.dsw_strerror(code)) > 5
    assert b"unknown" in lib.dsw_strerror(-99)


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    # argument validation happens before any CUDA call, so these are safe on a CPU-only box
    assert lib.dsw_cheb_fwd_workspace_bytes(0, 10, 4, 4, 3) == 0
    assert lib.dsw_cheb_fwd(None, None, 0, 0, None, None, None, 1, 1, 1, 1, 0, None, 0, None) < 0
    assert lib.dsw_spmm_fwd(None, None, 0, 0, None, 1, 1, None) < 0
    handle = C.c_void_p()
    assert lib.dsw_plan_create(0, 0, 0, None, None, None, None, C.byref(handle)) < 0
    assert not handle.value


def test_new_entry_points_validate_arguments_without_a_gpu(lib):
    assert lib.dsw_rezero_bwd_workspace_bytes() >= 4
    assert lib.dsw_rezero_fwd(None, None, None, None, 16, None) < 0
    assert lib.dsw_rezero_bwd(None, None, None, None, None, None, 0, 16, None) < 0
    assert lib.dsw_linear_rezero_fwd(None, 0, 0, None, None, None, None, None, 1, 1, 4, 4, None, 0, None) < 0
    assert lib.dsw_debug_dense_counters(None, 0) < 0
    # strided / accumulating forms of the skip-connection path
    assert lib.dsw_linear_rezero_fwd_ld(None, 0, 0, None, None, None, None, None, 8, 1, 1, 4, 4, None, 0, None) < 0
    assert lib.dsw_linear_bwd_acc(None, 0, 0, None, None, None, None, None, None, 1, 1, 4, 4, None, 0, None) < 0
    assert lib.dsw_spmm_fwd_ex(None, None, 0, 0, None, 0, 0, None, 0, 0, 1, 4, None) < 0
    assert lib.dsw_spmm_bwd_ex(None, None, 0, 0, None, 0, 0, None, 0, 0, 1, 4, None) < 0


def test_tuning_options_round_trip_and_env_hook(lib):
    """dsw_set_option / dsw_get_option and the DSW_OPTIONS="key=value,..." hook of _lib.load()."""
    import re
    import subprocess

    with open(_lib.HEADER_PATH) as f:
        count = int(re.search(r"DSW_OPT_COUNT\s*=\s*(\d+)", f.read()).group(1))
    for key in range(count):
        prev = lib.dsw_get_option(key)
        assert lib.dsw_set_option(key, 3) == 0 and lib.dsw_get_option(key) == 3
        assert lib.dsw_set_option(key, prev) == 0
    assert lib.dsw_set_option(count, 1) < 0 and lib.dsw_set_option(0, -1) < 0 and lib.dsw_get_option(count) == -1
    code = ("import sys; sys.path.insert(0, %r); from deepsphere_weather_b200 import _lib; l = _lib.load(); "
            "print(l.dsw_get_option(15), l.dsw_get_option(14))" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DSW_OPTIONS="15=1,14=2"), capture_output=True, text=True)
    assert out.stdout.split() == ["1", "2"], out.stderr[-400:]
    bad = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DSW_OPTIONS="999=1"), capture_output=True, text=True)
    assert bad.returncode != 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from deepsphere_weather_b200 import functional as F_
    from deepsphere_weather_b200 import graphs as G
    from deepsphere_weather_b200 import layers as L

    lap = G.healpix_laplacian(2)
    layer = L.ConvCheb(4, 4, 2, lap)
    with pytest.raises(RuntimeError, match="no CPU"):
        layer(torch.randn(1, 48, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        F_.plan_for(lap)


def test_shard_batch_partitions_exactly():
    for total in (1, 7, 32, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_batch(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a1 >= a0 and b1 >= b0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ddp_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)  # different init per rank: the bucket must broadcast rank 0's
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        bucket = FlatGradBucket(net)
        g = torch.Generator().manual_seed(7)
        x, y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
        s0, s1 = shard_batch(8, rank, world)
        bucket.zero_()
        # sum-of-squares / global batch so that the mean over ranks of shard grads * world == full grad
        loss = ((net(x[s0:s1]) - y[s0:s1]) ** 2).sum() / 8 * world
        loss.backward()
        bucket.check_views()
        bucket.allreduce_mean()
        if rank == 0:
            ret["flat"] = bucket.flat.clone()
            ret["params"] = [p.detach().clone() for p in net.parameters()]
            ret["x"], ret["y"] = x, y
    finally:
        dist.destroy_process_group()


def test_flat_bucket_allreduce_equals_single_process_gradient_gloo_ws2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ddp_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    with torch.no_grad():
        for p, q in zip(net.parameters(), ret["params"]):
            p.copy_(q)
    loss = ((net(ret["x"]) - ret["y"]) ** 2).sum() / 8
    loss.backward()
    full = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(ret["flat"], full, rtol=1e-5, atol=1e-6)


def _ddp_uneven_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        bucket = FlatGradBucket(net)
        g = torch.Generator().manual_seed(7)
        total = 7  # 4 + 3 samples: uneven shards
        x, y = torch.randn(total, 6, generator=g), torch.randn(total, 3, generator=g)
        s0, s1 = shard_batch(total, rank, world)
        bucket.zero_()
        loss = ((net(x[s0:s1]) - y[s0:s1]) ** 2).mean()   # a per-shard MEAN, as the training loop computes it
        loss.backward()
        bucket.allreduce_mean(local_batch=s1 - s0, global_batch=total)
        if rank == 0:
            ret["flat"] = bucket.flat.clone()
            ret["params"] = [p.detach().clone() for p in net.parameters()]
            ret["x"], ret["y"] = x, y
    finally:
        dist.destroy_process_group()


def test_uneven_shards_are_weighted_by_their_share_of_the_batch_gloo_ws2():
    """a mean of per-shard means over-weights the smaller shard; weighted by local / global batch the
    all-reduced gradient equals the single-process global-batch mean gradient."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ddp_uneven_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    with torch.no_grad():
        for p, q in zip(net.parameters(), ret["params"]):
            p.copy_(q)
    loss = ((net(ret["x"]) - ret["y"]) ** 2).mean()
    loss.backward()
    full = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(ret["flat"], full, rtol=1e-5, atol=1e-6)


def test_lmax_estimate_is_converged_and_deterministic():
    """the estimate must not undershoot (the rescaled Laplacian has to stay inside [-1, 1])."""
    import numpy as np
    from scipy.sparse import linalg as sla

    from deepsphere_weather_b200 import graphs as G

    for nside in (4, 16):
        L = G.knn_laplacian(G.healpix_nested_xyz(nside), 20)
        a, b = G.estimate_lmax_deterministic(L), G.estimate_lmax_deterministic(L)
        assert a == b
        true = float(sla.eigsh(L.astype(np.float64), k=1, which="LA", return_eigenvectors=False, tol=1e-10)[0])
        assert true <= a <= true * 1.011, (nside, a, true)
        Ls = G.prepare_torch_laplacian(L).to_dense().double().numpy()
        ev = np.linalg.eigvalsh(Ls) if nside == 4 else None
        if ev is not None:
            assert ev.max() <= 1.0 + 1e-6 and ev.min() >= -1.0 - 1e-6
