"""TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference hot path.

Nothing in the product package imports this module.  It may be imported by ``tests/``, by
``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs,
and only as the checker / the reported CPU baseline.

Parity status: **pinned against the reference run in the build container** — the reference ships
no tests or golden vectors of its own, so ``oracle/make_golden.py`` imports
the unmodified reference (``oracle/ref_import.py``), runs it on seeded inputs and commits the
input/output vectors to ``tests/golden/``; ``tests/test_oracle_golden.py`` holds this file to
those vectors.

Every function cites the reference lines it restates (paths relative to ``/root/reference``).
The arithmetic is fp32 on CPU torch (the reference's own arithmetic is ATen ``sparse.mm`` /
``matmul``); ``*_dense_f64`` is an independent numpy fp64 restatement used as a second opinion.
"""
from __future__ import annotations

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# Chebyshev convolution                                            modules/layers.py:113-180
# --------------------------------------------------------------------------------------


def conv_cheb(laplacian: torch.Tensor, inputs: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """``y[b,v,:] = sum_k (T_k(L) x_b)[v,:] @ W[:,k,:]`` — restates ``layers.py:141-178``.

    Same operation order as the reference: the batch is folded into the SpMM's column dimension
    (``[V, Fin*B]``, ``layers.py:158-159``), the recurrence is ``x1 = L x0``,
    ``x_k = 2 L x_{k-1} - x_{k-2}`` (``:163-169``), and the K terms are mixed by one dense matmul
    with ``W`` viewed as ``[Fin*K, Fout]`` (``:171-177``) — i.e. reduction index ``fin*K + k``.
    """
    n_b, n_v, f_in = inputs.shape
    f_in_w, n_k, f_out = weight.shape
    if f_in != f_in_w:  # layers.py:149-154
        raise ValueError(
            "Input tensor shape does not match the expected shape: \n"
            f"- Input tensor shape :{f_in} \n- Expected tensor shape :{f_in_w} \n"
        )
    t_prev = inputs.permute(1, 2, 0).contiguous().view(n_v, f_in * n_b)
    stack = t_prev.unsqueeze(0)
    # the reference grows the stack with one torch.cat per term (layers.py:164-169: O(K^2) copies); kept as is so
    # that the CPU arm of bench.py pays what the reference pays
    if n_k > 1:
        t_cur = torch.sparse.mm(laplacian, t_prev)
        stack = torch.cat((stack, t_cur.unsqueeze(0)), 0)
        for _ in range(2, n_k):
            t_next = 2 * torch.sparse.mm(laplacian, t_cur) - t_prev
            stack = torch.cat((stack, t_next.unsqueeze(0)), 0)
            t_prev, t_cur = t_cur, t_next
    stack = stack.view(n_k, n_v, f_in, n_b)
    stack = stack.permute(3, 1, 2, 0).contiguous().view(n_b * n_v, f_in * n_k)
    out = stack.matmul(weight.view(f_in * n_k, f_out))
    return out.view(n_b, n_v, f_out)


def conv_cheb_layer(laplacian, inputs, weight, bias=None):
    """``ConvCheb.forward`` (``layers.py:365-376``): convolution, then bias added in place."""
    out = conv_cheb(laplacian, inputs, weight)
    if bias is not None:
        out += bias
    return out


def conv_cheb_dense_f64(lap_dense: np.ndarray, x: np.ndarray, w: np.ndarray, bias=None) -> np.ndarray:
    """Independent fp64 restatement: explicit dense ``T_k(L)`` matrices, then one contraction."""
    lap = np.asarray(lap_dense, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    n_k = w.shape[1]
    t = [np.eye(lap.shape[0])]
    if n_k > 1:
        t.append(lap)
    for _ in range(2, n_k):
        t.append(2.0 * lap @ t[-1] - t[-2])
    tk = np.stack(t[:n_k], 0)  # [K, V, V]
    y = np.einsum("kvu,buf,fko->bvo", tk, x, w, optimize=True)
    if bias is not None:
        y = y + np.asarray(bias, dtype=np.float64)
    return y


def he_normal_std(in_channels: int, kernel_size: int) -> float:
    """``ConvCheb.reset_parameters`` default (``layers.py:291-335``): relu / fan-in / normal."""
    return float(np.sqrt(2.0 / (in_channels * kernel_size)))


# --------------------------------------------------------------------------------------
# Sparse remap pooling / unpooling                                modules/layers.py:948-1036
# --------------------------------------------------------------------------------------


def remap(matrix: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """``RemapBlock.forward`` (``layers.py:956-964``): ``out[b,v',f] = sum_v M[v',v] x[b,v,f]``.

    Returns the same non-contiguous ``[B, V', F]`` view (strides ``(1, F*B, B)``) as the reference.
    """
    n_b, n_v, n_f = x.shape
    n_new = matrix.shape[0]
    flat = x.permute(1, 2, 0).reshape(n_v, n_f * n_b)
    flat = torch.sparse.mm(matrix, flat)
    return flat.reshape(n_new, n_f, n_b).permute(2, 0, 1)


def max_area_pool_matrix(mat_csr) -> torch.Tensor:
    """``GeneralMaxAreaPool.process_remap_matrix`` (``layers.py:1000-1015``): one 1.0 per coarse
    row, at the column holding the row's largest weight (first on ties, ``np.argmax``)."""
    dense_arg = np.asarray(np.argmax(mat_csr, axis=1)).ravel().astype(np.int64)
    rows = np.arange(dense_arg.size, dtype=np.int64)
    idx = torch.from_numpy(np.stack([rows, dense_arg]))
    return torch.sparse_coo_tensor(idx, torch.ones(rows.size), mat_csr.shape, dtype=torch.float32).coalesce()


def max_area_unpool_matrix(mat_csr) -> torch.Tensor:
    """``GeneralMaxAreaUnpool.process_remap_matrix`` (``layers.py:1021-1036``): the argument is
    ``pool.T`` (fine x coarse); for every *column* keep a single 1.0 at the row of its maximum."""
    arg = np.asarray(np.argmax(mat_csr, axis=0)).ravel().astype(np.int64)
    cols = np.arange(arg.size, dtype=np.int64)
    idx = torch.from_numpy(np.stack([arg, cols]))
    return torch.sparse_coo_tensor(idx, torch.ones(cols.size), mat_csr.shape, dtype=torch.float32).coalesce()


# --------------------------------------------------------------------------------------
# Max-value pooling with index output                            modules/layers.py:1040-1103
# --------------------------------------------------------------------------------------


def maxval_pool(matrix: torch.Tensor, x: torch.Tensor):
    """``GeneralMaxValPool.forward`` (``layers.py:1043-1083``).

    For coarse row ``r`` and column ``c = f*B + b`` pick ``j* = argmax_j M[r,j] * x[j,c]`` over the
    row's stored entries (first maximum wins, like ``torch.argmax``), output the *unweighted*
    ``x[j*, c]`` and the int64 index pairs ``(j*, c)`` laid out ``[2, (F*B) * V']`` with ``c``
    major (``:1075-1079``).  The reference loops over rows in Python; this is the same selection
    done with a segmented scan.
    """
    n_b, n_v, n_f = x.shape
    n_new = matrix.shape[0]
    m = matrix.coalesce()
    row, col = m.indices()
    wts = m.values()
    flat = x.permute(1, 2, 0).reshape(n_v, n_f * n_b)
    n_c = flat.shape[1]
    counts = torch.bincount(row, minlength=n_new)
    assert int(counts.min()) > 0, "every coarse row needs at least one entry"
    starts = torch.cumsum(counts, 0) - counts
    picked = torch.empty(n_new, n_c, dtype=torch.int64)
    flat_d = flat.detach()
    for r in range(n_new):  # segments are short (4-9 entries); only the argmax is per-row
        s, e = int(starts[r]), int(starts[r] + counts[r])
        cand = wts[s:e, None] * flat_d[col[s:e]]
        picked[r] = col[s:e][torch.argmax(cand, dim=0)]
    pooled = torch.gather(flat, 0, picked)
    col_ids = torch.arange(n_c, dtype=torch.int64).expand(n_new, n_c)
    index = torch.stack([picked, col_ids], dim=2).permute(1, 0, 2).reshape(-1, 2).T
    return pooled.reshape(n_new, n_f, n_b).permute(2, 0, 1), index


def maxval_unpool(n_fine: int, x: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """``GeneralMaxValUnpool.forward`` (``layers.py:1089-1103``): zeros ``[V, F*B]`` and one
    ``index_put`` of ``x`` flattened in ``(f, b, v')`` order."""
    n_b, _, n_f = x.shape
    vals = x.permute(2, 0, 1).flatten()
    out = torch.zeros(n_fine, n_b * n_f, dtype=x.dtype)
    out = torch.index_put(out, (index[0], index[1]), vals)
    return out.reshape(n_fine, n_f, n_b).permute(2, 0, 1)


# --------------------------------------------------------------------------------------
# Nested-order HEALPix pools                                       modules/layers.py:784-941
# --------------------------------------------------------------------------------------


def healpix_max_pool(x: torch.Tensor, kernel: int = 4):
    """``HealpixMaxPool.forward`` (``layers.py:800-830``): windows of ``kernel`` consecutive
    nested pixels; indices are int64 ``[B, F, V/kernel]`` positions along the fine node axis."""
    b, v, f = x.shape
    win = x.permute(0, 2, 1).reshape(b, f, v // kernel, kernel)
    arg = torch.argmax(win, dim=3)  # first max on ties, NaN counts as max — same as max_pool1d
    val = torch.gather(win, 3, arg.unsqueeze(3)).squeeze(3)
    idx = arg + torch.arange(v // kernel, dtype=torch.int64) * kernel
    return val.permute(0, 2, 1), idx


def healpix_max_unpool(x: torch.Tensor, indices: torch.Tensor, kernel: int = 4) -> torch.Tensor:
    """``HealpixMaxUnpool.forward`` (``layers.py:843-863``)."""
    b, vc, f = x.shape
    out = torch.zeros(b, f, vc * kernel, dtype=x.dtype)
    out.scatter_(2, indices, x.permute(0, 2, 1))
    return out.permute(0, 2, 1)


def healpix_avg_pool(x: torch.Tensor, kernel: int = 4):
    """``HealpixAvgPool.forward`` (``layers.py:883-900``); second output is ``None``."""
    b, v, f = x.shape
    win = x.permute(0, 2, 1).reshape(b, f, v // kernel, kernel)
    acc = win[..., 0].clone()
    for i in range(1, kernel):
        acc = acc + win[..., i]
    return (acc / kernel).permute(0, 2, 1), None


def healpix_avg_unpool(x: torch.Tensor, kernel: int = 4) -> torch.Tensor:
    """``HealpixAvgUnpool.forward`` (``layers.py:921-941``): nearest-neighbour repeat."""
    return x.repeat_interleave(kernel, dim=1)


# --------------------------------------------------------------------------------------
# nn.Module wrappers so that the whole U-Net can be run on the oracle (CPU baseline leg)
# --------------------------------------------------------------------------------------


class OracleConvCheb(torch.nn.Module):
    """Same constructor / parameters / buffer as ``ConvCheb`` (``layers.py:223-251``)."""

    def __init__(self, in_channels, out_channels, kernel_size, laplacian, bias=True, **_):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.register_buffer("laplacian", laplacian)
        self.weight = torch.nn.Parameter(torch.empty(in_channels, kernel_size, out_channels))
        self.bias = torch.nn.Parameter(torch.zeros(out_channels)) if bias else None
        torch.nn.init.normal_(self.weight, 0.0, he_normal_std(in_channels, kernel_size))

    def forward(self, inputs):
        return conv_cheb_layer(self.laplacian, inputs, self.weight, self.bias)


class OracleRemap(torch.nn.Module):
    def __init__(self, matrix: torch.Tensor, returns_index: bool):
        super().__init__()
        self.register_buffer("remap_matrix", matrix)
        self.returns_index = returns_index

    def forward(self, x, *args, **kwargs):
        out = remap(self.remap_matrix, x)
        return (out, None) if self.returns_index else out


class OracleHealpixPool(torch.nn.Module):
    def __init__(self, mode: str, kernel_size: int = 4):
        super().__init__()
        self.mode, self.kernel_size = mode, kernel_size

    def forward(self, x):
        if self.mode == "max":
            return healpix_max_pool(x, self.kernel_size)
        return healpix_avg_pool(x, self.kernel_size)


class OracleHealpixUnpool(torch.nn.Module):
    def __init__(self, mode: str, kernel_size: int = 4):
        super().__init__()
        self.mode, self.kernel_size = mode, kernel_size

    def forward(self, x, indices=None, *args):
        if self.mode == "max":
            return healpix_max_unpool(x, indices, self.kernel_size)
        return healpix_avg_unpool(x, self.kernel_size)
