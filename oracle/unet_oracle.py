"""TEST INFRASTRUCTURE ONLY — the U-Net architecture instantiated on the oracle's CPU layers.

Builds ``deepsphere_weather_b200.models.UNetSpherical`` (pure composition code, validated against
the unmodified reference by ``tests/test_oracle_golden.py``) with the oracle layer classes from
``oracle/cheb_oracle.py`` injected, so the whole forward+backward runs on CPU torch through the
reference's arithmetic (``torch.sparse.mm`` / ``matmul`` / pool restatements).  Used as the
checker for whole-model parity and as the reported CPU baseline (``bench.py``).
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch
from scipy import sparse

from . import cheb_oracle as O


def _coo(mat) -> torch.Tensor:
    m = sparse.coo_matrix(mat)
    idx = torch.from_numpy(np.stack([m.row.astype(np.int64), m.col.astype(np.int64)]))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(m.data.astype(np.float32)), m.shape,
                                   check_invariants=False).coalesce()


class _MaxValPool(torch.nn.Module):
    def __init__(self, mat):
        super().__init__()
        self.register_buffer("remap_matrix", _coo(mat))

    def forward(self, x, *a, **k):
        return O.maxval_pool(self.remap_matrix, x)


class _MaxValUnpool(torch.nn.Module):
    def __init__(self, mat):
        super().__init__()
        self.register_buffer("remap_matrix", _coo(mat))

    def forward(self, x, index, *a, **k):
        return O.maxval_unpool(self.remap_matrix.shape[0], x, index)


def _general_pools(pool_method: str, matrices, **_):
    pool_mat, unpool_mat = matrices
    if pool_method == "interp":
        return O.OracleRemap(_coo(pool_mat), True), O.OracleRemap(_coo(unpool_mat), False)
    if pool_method == "maxarea":
        return (O.OracleRemap(O.max_area_pool_matrix(sparse.csr_matrix(pool_mat)), True),
                O.OracleRemap(O.max_area_unpool_matrix(sparse.csr_matrix(sparse.coo_matrix(pool_mat).T)), False))
    if pool_method == "maxval":
        return _MaxValPool(pool_mat), _MaxValUnpool(unpool_mat)
    raise ValueError(pool_method)


class _HPool(O.OracleHealpixPool):
    def __init__(self, mode, kernel_size=4):
        super().__init__(mode, kernel_size)


def oracle_backend() -> SimpleNamespace:
    mk = lambda mode: (lambda kernel_size=4, **_: O.OracleHealpixPool(mode, kernel_size),
                       lambda kernel_size=4, **_: O.OracleHealpixUnpool(mode, kernel_size))
    return SimpleNamespace(ConvCheb=O.OracleConvCheb, healpix_pools={"max": mk("max"), "avg": mk("avg")},
                           general_pools=_general_pools)


def build_unet_oracle(*args, **kwargs):
    from deepsphere_weather_b200.models import UNetSpherical

    return UNetSpherical(*args, backend=oracle_backend(), **kwargs)


def fill_parameters(model: torch.nn.Module, seed: int = 0, rezero: float = 1.0) -> None:
    """Deterministic parameter fill shared by every parity harness (see
    ``deepsphere_weather_b200.models.deterministic_fill``)."""
    from deepsphere_weather_b200.models import deterministic_fill

    deterministic_fill(model, seed, rezero)
