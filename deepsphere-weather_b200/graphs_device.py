"""Device-side construction of the sparse operators that feed the hot path.

The reference builds its graphs with pygsp on the host (``modules/models.py:43-46``), estimates the largest eigenvalue
with a randomly started ARPACK run (``modules/layers.py:57-69``) and gets its pooling weights from the CDO binary
(``modules/layers.py:531-581``).  Here the same operators are produced by CUDA kernels (``csrc/dsw_graph.cu``) straight
into device memory, deterministically: brute-force fp64 k-NN on the unit vectors, Gaussian weights, max-symmetrisation,
normalised Laplacian, a converged power iteration for ``lmax``, the rescaling ``2 L / lmax - I``, and the exact nested
pool / unpool matrices.  Only the pixel centres (``graphs.healpix_nested_xyz``: O(V) numpy) are computed on the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from . import graphs as G
from .functional import _stream_ptr


def knn_laplacian_device(xyz, k: int = 20, device="cuda", rescale: bool = True, lmax: float | None = None):
    """Coalesced sparse-COO (rescaled) Laplacian of the symmetrised Gaussian ``k``-NN graph of the points ``xyz``
    (``[V, 3]`` unit vectors, numpy or tensor), built on ``device``.  Returns ``(laplacian, lmax)``."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("knn_laplacian_device builds on a CUDA device (graphs.knn_laplacian is the host builder)")
    pts = torch.as_tensor(np.asarray(xyz, dtype=np.float64) if not torch.is_tensor(xyz) else xyz, dtype=torch.float64).to(device).contiguous()
    V = int(pts.shape[0])
    k = min(int(k), V - 1)
    lib = _lib.load()
    cap = int(lib.dsw_graph_nnz_capacity(V, k))
    rows = torch.empty(cap, dtype=torch.int64, device=device)
    cols = torch.empty(cap, dtype=torch.int64, device=device)
    vals = torch.empty(cap, dtype=torch.float32, device=device)
    ws = torch.empty(max(int(lib.dsw_graph_workspace_bytes(V, k)), 1), dtype=torch.uint8, device=device)
    nnz, lmax_out = C.c_int64(0), C.c_double(0.0)
    with torch.cuda.device(device):
        rc = lib.dsw_graph_knn_laplacian(pts.data_ptr(), V, k, 1 if rescale else 0, float(lmax) if lmax else 0.0, cap, rows.data_ptr(),
                                         cols.data_ptr(), vals.data_ptr(), C.byref(nnz), C.byref(lmax_out), ws.data_ptr(), ws.numel(),
                                         _stream_ptr(device))
    _lib.check(rc, "dsw_graph_knn_laplacian")
    n = int(nnz.value)
    lap = torch.sparse_coo_tensor(torch.stack([rows[:n], cols[:n]]), vals[:n], (V, V), check_invariants=False, is_coalesced=True)
    return lap, float(lmax_out.value)


def healpix_laplacian_device(nside: int, k: int = 20, device="cuda") -> torch.Tensor:
    """``graphs.healpix_laplacian`` built on the device (nested ordering)."""
    return knn_laplacian_device(G.healpix_nested_xyz(nside), k, device)[0]


def nested_pool_matrices_device(n_fine: int, kernel: int = 4, device="cuda"):
    """``graphs.nested_pool_matrices`` as coalesced torch COO tensors on the device: ``(pool [V/kernel, V], unpool [V, V/kernel])``."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("nested_pool_matrices_device builds on a CUDA device")
    lib = _lib.load()
    idx = [torch.empty(n_fine, dtype=torch.int64, device=device) for _ in range(4)]
    pv, uv = (torch.empty(n_fine, dtype=torch.float32, device=device) for _ in range(2))
    with torch.cuda.device(device):
        rc = lib.dsw_graph_nested_pool(n_fine, kernel, idx[0].data_ptr(), idx[1].data_ptr(), pv.data_ptr(), idx[2].data_ptr(),
                                       idx[3].data_ptr(), uv.data_ptr(), _stream_ptr(device))
    _lib.check(rc, "dsw_graph_nested_pool")
    n_coarse = n_fine // kernel
    pool = torch.sparse_coo_tensor(torch.stack(idx[:2]), pv, (n_coarse, n_fine), check_invariants=False, is_coalesced=True)
    unpool = torch.sparse_coo_tensor(torch.stack(idx[2:]), uv, (n_fine, n_coarse), check_invariants=False, is_coalesced=True)
    return pool, unpool
