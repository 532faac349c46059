"""Host-side construction of the sparse operators that are *inputs* to the hot path.

The reference obtains its graphs from ``pygsp`` (``modules/models.py:43-46``) and its pooling
matrices from ``xsphere``/CDO (``modules/layers.py:531-581``); neither is installable here, and
SURVEY.md §8c treats both as inputs.  This module builds equivalent synthetic operators with
numpy/scipy only:

* HEALPix pixel centres in NESTED order (the standard ``pix2vec`` arithmetic, restated);
* equiangular (lat x lon) grids;
* k-NN graph, Gaussian kernel weights, max-symmetrisation, normalised Laplacian;
* rescaling ``2 L / lmax - I`` (what ``prepare_torch_laplacian`` does at
  ``modules/layers.py:82-106``) with a *deterministic* largest-eigenvalue estimate;
* nested 4-children pool / unpool matrices and random "Voronoi-like" row-stochastic ones.

Nothing here runs on the GPU; it is model-construction code executed once.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy import sparse
from scipy.spatial import cKDTree

_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4], dtype=np.int64)
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7], dtype=np.int64)


def _compact_bits(v: np.ndarray) -> np.ndarray:
    """Gather the even-position bits of ``v`` into a dense integer (inverse of bit spreading)."""
    v = v & 0x5555555555555555
    v = (v | (v >> 1)) & 0x3333333333333333
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0F
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FF
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFF
    v = (v | (v >> 16)) & 0x00000000FFFFFFFF
    return v


def healpix_nested_lonlat(nside: int):
    """Longitude/colatitude (radians) of every HEALPix pixel centre, NESTED ordering.

    ``nside`` must be a power of two.  Returns ``(theta, phi)`` with ``theta`` the colatitude.
    """
    if nside < 1 or (nside & (nside - 1)):
        raise ValueError("nside must be a power of two")
    npix = 12 * nside * nside
    p = np.arange(npix, dtype=np.int64)
    npface = nside * nside
    face = p // npface
    ipf = p % npface
    ix = _compact_bits(ipf)
    iy = _compact_bits(ipf >> 1)
    jr = _JRLL[face] * nside - ix - iy - 1  # ring number, 1 .. 4*nside-1
    nl4 = 4 * nside

    z = np.empty(npix, dtype=np.float64)
    nr = np.empty(npix, dtype=np.int64)
    kshift = np.zeros(npix, dtype=np.int64)

    north = jr < nside
    south = jr > 3 * nside
    equat = ~(north | south)

    nr[north] = jr[north]
    z[north] = 1.0 - nr[north].astype(np.float64) ** 2 / (3.0 * npface)
    nr[south] = nl4 - jr[south]
    z[south] = nr[south].astype(np.float64) ** 2 / (3.0 * npface) - 1.0
    nr[equat] = nside
    z[equat] = (2 * nside - jr[equat]) * (2.0 / (3.0 * nside))
    kshift[equat] = (jr[equat] - nside) & 1

    jp = (_JPLL[face] * nr + ix - iy + 1 + kshift) // 2
    jp = np.where(jp > nl4, jp - nl4, jp)
    jp = np.where(jp < 1, jp + nl4, jp)
    phi = (jp - (kshift + 1) * 0.5) * (np.pi / 2.0 / nr)
    theta = np.arccos(np.clip(z, -1.0, 1.0))
    return theta, phi


def sphere_xyz(theta: np.ndarray, phi: np.ndarray) -> np.ndarray:
    st = np.sin(theta)
    return np.stack([st * np.cos(phi), st * np.sin(phi), np.cos(theta)], axis=1)


def healpix_nested_xyz(nside: int) -> np.ndarray:
    return sphere_xyz(*healpix_nested_lonlat(nside))


def equiangular_xyz(nlat: int, nlon: int) -> np.ndarray:
    """Row-major (lat-major) equiangular grid, cell centres, poles excluded."""
    theta = (np.arange(nlat) + 0.5) * (np.pi / nlat)
    phi = np.arange(nlon) * (2.0 * np.pi / nlon)
    tt, pp = np.meshgrid(theta, phi, indexing="ij")
    return sphere_xyz(tt.ravel(), pp.ravel())


def knn_laplacian(xyz: np.ndarray, k: int = 20) -> sparse.csr_matrix:
    """Normalised Laplacian ``I - D^-1/2 W D^-1/2`` of the symmetrised Gaussian k-NN graph (fp64).

    Mirrors what the reference asks ``pygsp`` for (``lap_type="normalized"``, ``models.py:45``):
    weights ``exp(-d^2 / (2 sigma^2))`` with ``sigma`` the mean neighbour distance, symmetrised by
    ``max(W, W^T)``.
    """
    n = xyz.shape[0]
    k = min(k, n - 1)
    dist, idx = cKDTree(xyz).query(xyz, k=k + 1)
    dist, idx = dist[:, 1:], idx[:, 1:]
    sigma = dist.mean()
    w = np.exp(-(dist**2) / (2.0 * sigma**2))
    rows = np.repeat(np.arange(n), k)
    W = sparse.csr_matrix((w.ravel(), (rows, idx.ravel())), shape=(n, n))
    W = W.maximum(W.T).tocsr()
    d = np.asarray(W.sum(axis=1)).ravel()
    dinv = 1.0 / np.sqrt(d)
    L = sparse.identity(n, format="csr") - sparse.diags(dinv) @ W @ sparse.diags(dinv)
    L = L.tocsr()
    L.sort_indices()
    return L


def estimate_lmax_deterministic(L: sparse.spmatrix, iters: int = 64) -> float:
    """Largest eigenvalue of ``L`` with the reference's 1 % safety margin (``estimate_lmax``, ``layers.py:57-69``),
    deterministically.  The reference asks ARPACK with a *random* start vector, so two of its calls differ by ~1e-3
    (SURVEY.md §0.4); here ARPACK gets a fixed start vector and a tight tolerance, which makes the estimate reproducible
    AND converged — an unconverged power iteration approaches lmax from below and would leave the rescaled operator
    ``2 L / lmax - I`` with eigenvalues above 1, outside the interval the Chebyshev recurrence assumes.  Like the
    reference the result is capped at 2 (the bound for normalised Laplacians is exact there, ``layers.py:66-68``)."""
    from scipy.sparse import linalg as sla

    n = L.shape[0]
    v0 = np.cos(np.arange(n, dtype=np.float64) * 0.7390851332151607) + 1.5
    Ld = sparse.csr_matrix(L, dtype=np.float64)
    lam = None
    if n > 3:
        try:
            lam = float(sla.eigsh(Ld, k=1, which="LA", v0=v0, tol=1e-9, maxiter=max(10 * n, 1000), return_eigenvectors=False)[0])
        except sla.ArpackError:
            lam = None
    if lam is None:  # tiny operators / no convergence: power iteration until the Rayleigh quotient settles
        v = v0 / np.linalg.norm(v0)
        lam = 0.0
        for _ in range(max(iters, 1) * 64):
            w = Ld @ v
            nw = float(np.linalg.norm(w))
            if nw == 0.0:
                break
            new = float(v @ w)
            v = w / nw
            if abs(new - lam) <= 1e-10 * max(abs(new), 1.0):
                lam = new
                break
            lam = new
    return lam * (1.0 + 2.0 * 5e-3)


def scipy_to_torch_coo(mat: sparse.spmatrix, dtype=torch.float32) -> torch.Tensor:
    """scipy sparse -> coalesced ``torch.sparse_coo_tensor`` with int64 indices (the boundary type,
    ``layers.py:584-594``)."""
    coo = sparse.coo_matrix(mat)
    idx = torch.from_numpy(np.stack([coo.row.astype(np.int64), coo.col.astype(np.int64)]))
    val = torch.from_numpy(coo.data.astype(np.float32)).to(dtype)
    return torch.sparse_coo_tensor(idx, val, coo.shape, dtype=dtype, check_invariants=False).coalesce()


def prepare_torch_laplacian(L: sparse.spmatrix, lmax: float | None = None) -> torch.Tensor:
    """fp64 scipy Laplacian -> fp32, rescaled to ``2 L / lmax - I``, coalesced torch COO.

    Same contract as the reference's function of the same name (``layers.py:82-106``) except that
    ``lmax`` is deterministic (or caller-supplied)."""
    L = sparse.csr_matrix(L).astype(np.float32)
    if lmax is None:
        lmax = estimate_lmax_deterministic(L.astype(np.float64))
    L = L * np.float32(2.0 / lmax) - sparse.identity(L.shape[0], dtype=np.float32, format="csr")
    return scipy_to_torch_coo(L)


def healpix_laplacian(nside: int, k: int = 20) -> torch.Tensor:
    return prepare_torch_laplacian(knn_laplacian(healpix_nested_xyz(nside), k))


def equiangular_laplacian(nlat: int, nlon: int, k: int = 20) -> torch.Tensor:
    """Rescaled k-NN Laplacian of the row-major equiangular grid (BASELINE.json cfg5: 200 x 400)."""
    return prepare_torch_laplacian(knn_laplacian(equiangular_xyz(nlat, nlon), k))


def nested_pool_matrices(n_fine: int, kernel: int = 4):
    """Exact pool/unpool pair for nested orderings: coarse pixel ``i`` owns fine pixels
    ``kernel*i .. kernel*i+kernel-1``.  pool rows = ``[1/kernel]*kernel``, unpool rows = ``[1]``
    (``tutorials/interpolation_pooling.ipynb`` cell 16); ``pool @ unpool = I``."""
    n_coarse = n_fine // kernel
    rows = np.repeat(np.arange(n_coarse), kernel)
    cols = np.arange(n_coarse * kernel)
    pool = sparse.coo_matrix((np.full(cols.size, 1.0 / kernel), (rows, cols)), shape=(n_coarse, n_fine))
    unpool = sparse.coo_matrix((np.ones(cols.size), (cols, rows)), shape=(n_fine, n_coarse))
    return pool, unpool


def random_overlap_pool_matrices(n_fine: int, n_coarse: int, seed: int = 0, extra: int = 3):
    """Row-stochastic "Voronoi-like" pool matrix with 4-9 nnz per row whose supports overlap, and
    the matching column-normalised-transposed unpool matrix (same normalisations as
    ``build_pooling_matrices`` at ``layers.py:576-581``)."""
    rng = np.random.default_rng(seed)
    ratio = n_fine // n_coarse
    rows, cols, vals = [], [], []
    for r in range(n_coarse):
        base = np.arange(r * ratio, min((r + 1) * ratio, n_fine))
        n_extra = rng.integers(0, extra + 3)
        lo, hi = max(0, r * ratio - 3 * ratio), min(n_fine, (r + 1) * ratio + 3 * ratio)
        others = rng.integers(lo, hi, size=n_extra)
        c = np.unique(np.concatenate([base, others]))
        w = rng.uniform(0.05, 1.0, size=c.size)
        rows.append(np.full(c.size, r)), cols.append(c), vals.append(w)
    rows, cols, vals = map(np.concatenate, (rows, cols, vals))
    area = sparse.csr_matrix((vals, (rows, cols)), shape=(n_coarse, n_fine))
    # every fine node must be covered so that the unpool column normalisation is defined
    assert (np.asarray(area.sum(0)).ravel() > 0).all()
    pool = sparse.coo_matrix(area.multiply(1.0 / area.sum(1)))
    unpool = sparse.coo_matrix(area.multiply(1.0 / area.sum(0)).T)
    return pool, unpool
