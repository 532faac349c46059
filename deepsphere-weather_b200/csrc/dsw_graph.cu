// Device-side construction of the operators that feed the hot path:
//   * the symmetrised Gaussian k-NN graph of a point set on the sphere and its normalised Laplacian
//     L = I - D^-1/2 W D^-1/2  — what the reference asks pygsp for (modules/models.py:43-46, utils_models.py:11-20:
//     SphereHealpix(subdivisions, nest, k, lap_type="normalized").L);
//   * the largest eigenvalue and the rescaling 2 L / lmax - I of prepare_torch_laplacian (modules/layers.py:57-106), with
//     a deterministic, converged power iteration in place of ARPACK's randomly started estimate;
//   * the exact nested-pixel pool / unpool matrices (tutorials/interpolation_pooling.ipynb cell 16; the reference gets
//     its weights from CDO, modules/layers.py:531-581).
// Everything is fp64 until the final cast, like the reference's scipy pipeline; neighbour ties are broken by the lower
// node index so that the graph is reproducible (cKDTree's tie order is not).  Output: coalesced COO (row-major, ascending
// columns), int64 indices + fp32 values — the boundary type of the reference's module buffers (layers.py:584-594).
#include <algorithm>
#include <cmath>

#include <cub/device/device_radix_sort.cuh>

#include "dsw_internal.cuh"

namespace dsw {
namespace {

constexpr int KNN_MAX_K = 64;
constexpr int KNN_THREADS = 128;

// Neighbour list of one query, kept sorted ascending by (squared chord length, node index): a tie goes to the lower index
// whatever the order in which the candidates are visited.
struct KnnList {
  double bd[KNN_MAX_K];
  int32_t bi[KNN_MAX_K];
  double worst;
  int32_t worst_i;
  int32_t k;
  __device__ __forceinline__ void init(int32_t kk) {
    k = kk;
    for (int i = 0; i < k; ++i) bd[i] = 1e300, bi[i] = 0x7fffffff;
    worst = 1e300, worst_i = 0x7fffffff;
  }
  __device__ __forceinline__ void offer(double d2, int32_t c) {
    if (d2 < worst || (d2 == worst && c < worst_i)) {
      int p = k - 1;
      while (p > 0 && (bd[p - 1] > d2 || (bd[p - 1] == d2 && bi[p - 1] > c))) {
        bd[p] = bd[p - 1], bi[p] = bi[p - 1];
        --p;
      }
      bd[p] = d2, bi[p] = c;
      worst = bd[k - 1], worst_i = bi[k - 1];
    }
  }
};
// no fused multiply-add: the squared chord length is rounded exactly like numpy's ((a - b) ** 2).sum(-1), so that ties
// fall the same way on both sides of a parity test
__device__ __forceinline__ double chord2(double ax, double ay, double az, double bx, double by, double bz) {
  const double dx = ax - bx, dy = ay - by, dz = az - bz;
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// Uniform grid over [-1, 1]^3 with cells of edge >= h: every point within distance h of a query lies in the 27 cells
// around the query's cell, so the k nearest found there are exact as soon as the k-th of them is within h.
__device__ __forceinline__ int32_t cell_coord(double x, double inv_cell, int32_t G) {
  const int32_t c = (int32_t)floor((x + 1.0) * inv_cell);
  return min(max(c, 0), G - 1);
}
__global__ void cell_id_kernel(const double* __restrict__ xyz, int32_t V, double inv_cell, int32_t G, uint32_t* __restrict__ cell,
                               int32_t* __restrict__ ident) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  const int32_t cx = cell_coord(xyz[3 * (size_t)i], inv_cell, G), cy = cell_coord(xyz[3 * (size_t)i + 1], inv_cell, G),
                cz = cell_coord(xyz[3 * (size_t)i + 2], inv_cell, G);
  cell[i] = (uint32_t)((cz * G + cy) * G + cx);
  ident[i] = i;
}
__global__ void cell_start_kernel(const uint32_t* __restrict__ sorted_cell, int32_t V, int32_t* __restrict__ cell_start,
                                  int32_t* __restrict__ cell_end) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  const uint32_t c = sorted_cell[i];
  if (i == 0 || sorted_cell[i - 1] != c) cell_start[c] = i;
  if (i == V - 1 || sorted_cell[i + 1] != c) cell_end[c] = i + 1;
}
// One thread per query (queries taken in cell order: the threads of a warp walk the same cells).
__global__ void __launch_bounds__(KNN_THREADS) knn_grid_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ sorted_pt, int32_t V,
                                                                int32_t k, double inv_cell, int32_t G, double h2,
                                                                const int32_t* __restrict__ cell_start, const int32_t* __restrict__ cell_end,
                                                                int32_t* __restrict__ nbr, double* __restrict__ nd2, int32_t* __restrict__ redo,
                                                                int32_t* __restrict__ n_redo) {
  const int s = blockIdx.x * KNN_THREADS + threadIdx.x;
  if (s >= V) return;
  const int32_t q = sorted_pt[s];
  const double qx = xyz[3 * (size_t)q], qy = xyz[3 * (size_t)q + 1], qz = xyz[3 * (size_t)q + 2];
  const int32_t cx = cell_coord(qx, inv_cell, G), cy = cell_coord(qy, inv_cell, G), cz = cell_coord(qz, inv_cell, G);
  KnnList L;
  L.init(k);
  for (int32_t z = max(cz - 1, 0); z <= min(cz + 1, G - 1); ++z)
    for (int32_t y = max(cy - 1, 0); y <= min(cy + 1, G - 1); ++y)
      for (int32_t x = max(cx - 1, 0); x <= min(cx + 1, G - 1); ++x) {
        const int32_t c = (z * G + y) * G + x;
        for (int32_t p = cell_start[c]; p < cell_end[c]; ++p) {
          const int32_t cj = sorted_pt[p];
          if (cj == q) continue;
          L.offer(chord2(qx, qy, qz, xyz[3 * (size_t)cj], xyz[3 * (size_t)cj + 1], xyz[3 * (size_t)cj + 2]), cj);
        }
      }
  if (!(L.worst <= h2)) {  // the k-th neighbour may lie outside the 27 cells: exact search for this query
    redo[atomicAdd(n_redo, 1)] = q;
    return;
  }
  for (int i = 0; i < k; ++i) nbr[(size_t)q * k + i] = L.bi[i], nd2[(size_t)q * k + i] = L.bd[i];
}
// Exact search over all points for the queries the grid could not settle (one warp-sized block per query would be
// wasteful: these are few, one thread each).
__global__ void __launch_bounds__(KNN_THREADS) knn_brute_kernel(const double* __restrict__ xyz, int32_t V, int32_t k, const int32_t* __restrict__ redo,
                                                                 const int32_t* __restrict__ n_redo, int32_t* __restrict__ nbr,
                                                                 double* __restrict__ nd2) {
  const int s = blockIdx.x * KNN_THREADS + threadIdx.x;
  if (s >= *n_redo) return;
  const int32_t q = redo[s];
  const double qx = xyz[3 * (size_t)q], qy = xyz[3 * (size_t)q + 1], qz = xyz[3 * (size_t)q + 2];
  KnnList L;
  L.init(k);
  for (int32_t cj = 0; cj < V; ++cj) {
    if (cj == q) continue;
    L.offer(chord2(qx, qy, qz, xyz[3 * (size_t)cj], xyz[3 * (size_t)cj + 1], xyz[3 * (size_t)cj + 2]), cj);
  }
  for (int i = 0; i < k; ++i) nbr[(size_t)q * k + i] = L.bi[i], nd2[(size_t)q * k + i] = L.bd[i];
}

// Deterministic fp64 sum of sqrt(nd2) over all V * k entries: fixed partition, fixed tree.
__global__ void __launch_bounds__(256) dist_partial_kernel(const double* __restrict__ nd2, int64_t n, double* __restrict__ partial) {
  __shared__ double sh[256];
  double acc = 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) acc += sqrt(nd2[i]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void sigma_kernel(const double* __restrict__ partial, int nparts, int64_t n, double* __restrict__ sigma) {
  double acc = 0;
  for (int i = 0; i < nparts; ++i) acc += partial[i];
  *sigma = acc / (double)n;
}

// Symmetrisation W = max(W, W^T): distances are symmetric, so the maximum is the union of the two directed edge sets.
// Pass 1 counts, per node j, the edges i -> j whose reverse j -> i is missing.
__device__ __forceinline__ bool has_neighbour(const int32_t* __restrict__ nbr, int32_t k, int32_t row, int32_t target) {
  for (int i = 0; i < k; ++i)
    if (nbr[(size_t)row * k + i] == target) return true;
  return false;
}
__global__ void count_missing_kernel(const int32_t* __restrict__ nbr, int32_t V, int32_t k, int32_t* __restrict__ extra) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)V * k) return;
  const int32_t i = (int32_t)(e / k), j = nbr[e];
  if (!has_neighbour(nbr, k, j, i)) atomicAdd(extra + j, 1);
}
// rowptr[i] = sum_{r < i} (k + extra[r] + 1)  (the + 1 is the diagonal); one block, fixed order.
__global__ void __launch_bounds__(1024) rowptr_kernel(const int32_t* __restrict__ extra, int32_t V, int32_t k, int64_t* __restrict__ rowptr) {
  __shared__ int64_t carry;
  __shared__ int64_t sh[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < V; base += 1024) {
    const int i = base + threadIdx.x;
    const int64_t mine = i < V ? (int64_t)k + extra[i] + 1 : 0;
    sh[threadIdx.x] = mine;
    __syncthreads();
    for (int s = 1; s < 1024; s <<= 1) {  // inclusive Hillis-Steele scan
      int64_t add = threadIdx.x >= s ? sh[threadIdx.x - s] : 0;
      __syncthreads();
      sh[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < V) rowptr[i] = carry + sh[threadIdx.x] - mine;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) rowptr[V] = carry;
}
// Pass 2 fills every row: the diagonal, the node's own k neighbours, and (through a per-row cursor) the reverse edges.
__global__ void fill_kernel(const int32_t* __restrict__ nbr, const double* __restrict__ nd2, int32_t V, int32_t k,
                            const int64_t* __restrict__ rowptr, int32_t* __restrict__ cursor, int32_t* __restrict__ col,
                            double* __restrict__ d2) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)V * k) return;
  const int32_t i = (int32_t)(e / k), s = (int32_t)(e - (int64_t)i * k), j = nbr[e];
  const int64_t at = rowptr[i] + 1 + s;
  col[at] = j, d2[at] = nd2[e];
  if (s == 0) col[rowptr[i]] = i, d2[rowptr[i]] = -1.0;  // diagonal marker
  if (!has_neighbour(nbr, k, j, i)) {
    const int32_t slot = atomicAdd(cursor + j, 1);
    const int64_t to = rowptr[j] + 1 + k + slot;
    col[to] = i, d2[to] = nd2[e];
  }
}
// Sort every row by column (rows are short) and turn distances into Gaussian weights; also the degree d_i = sum_j W_ij
// accumulated in ascending column order (deterministic).
__global__ void sort_weight_kernel(int32_t V, const int64_t* __restrict__ rowptr, int32_t* __restrict__ col, double* __restrict__ w,
                                   const double* __restrict__ sigma, double* __restrict__ dinv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  const int64_t a = rowptr[i], b = rowptr[i + 1];
  for (int64_t p = a + 1; p < b; ++p) {  // insertion sort by column
    const int32_t c = col[p];
    const double v = w[p];
    int64_t t = p;
    while (t > a && col[t - 1] > c) {
      col[t] = col[t - 1], w[t] = w[t - 1];
      --t;
    }
    col[t] = c, w[t] = v;
  }
  const double s2 = 2.0 * (*sigma) * (*sigma);
  double deg = 0;
  for (int64_t p = a; p < b; ++p) {
    if (col[p] == i) {
      w[p] = 0.0;
    } else {
      w[p] = exp(-w[p] / s2);
      deg += w[p];
    }
  }
  dinv[i] = 1.0 / sqrt(deg);
}
// L = I - (D^-1/2 W) D^-1/2 in the evaluation order of `sparse.diags(dinv) @ W @ sparse.diags(dinv)`.
__global__ void laplacian_kernel(int32_t V, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, double* __restrict__ w,
                                 const double* __restrict__ dinv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
    const int32_t c = col[p];
    w[p] = (c == i) ? 1.0 : -((dinv[i] * w[p]) * dinv[c]);
  }
}

// Largest eigenvalue by power iteration in ONE persistent launch: one block per SM (all co-resident), a counter-based
// grid barrier between the two phases of an iteration.  Fixed start vector, fixed partition of the rows, fixed
// reduction trees (per block, then over the blocks in index order): reproducible bit for bit.  Stops when the Rayleigh
// quotient has moved by less than `tol` (relative) over 8 iterations, or after `max_iter`.
constexpr int PI_THREADS = 256;
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int nblocks, unsigned int* gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    *gen += 1;
    __threadfence();
    atomicAdd(ctr, 1u);
    const unsigned int target = *gen * nblocks;
    while (*reinterpret_cast<volatile unsigned int*>(ctr) < target) __nanosleep(20);
    __threadfence();
  }
  __syncthreads();
}
__global__ void __launch_bounds__(PI_THREADS) power_iteration_kernel(int32_t V, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                                      const double* __restrict__ val, double* v, double* wv,
                                                                      double* partial /* [2][gridDim.x] */, unsigned int* ctr, int32_t max_iter,
                                                                      double tol, double* __restrict__ out) {
  __shared__ double red[2][PI_THREADS];
  __shared__ double lam_hist[8];
  __shared__ unsigned int gen;
  const int t = threadIdx.x, nb = gridDim.x;
  const int gtid = blockIdx.x * PI_THREADS + t, gstride = nb * PI_THREADS;
  if (t == 0) gen = 0;
  auto block_sum2 = [&](double a, double b) {  // block-wide sums of (a, b) -> partial[0][block], partial[1][block]
    red[0][t] = a, red[1][t] = b;
    __syncthreads();
    for (int s = PI_THREADS / 2; s > 0; s >>= 1) {
      if (t < s) red[0][t] += red[0][t + s], red[1][t] += red[1][t + s];
      __syncthreads();
    }
    if (t == 0) partial[blockIdx.x] = red[0][0], partial[nb + blockIdx.x] = red[1][0];
  };
  auto total2 = [&](double& a, double& b) {  // every thread sums the block partials in index order
    a = 0, b = 0;
    for (int i = 0; i < nb; ++i) a += __ldcg(partial + i), b += __ldcg(partial + nb + i);
  };
  {
    double acc = 0;
    for (int i = gtid; i < V; i += gstride) {
      const double x = cos((double)i * 0.7390851332151607) + 1.5;
      v[i] = x;
      acc += x * x;
    }
    block_sum2(acc, 0.0);
    grid_barrier(ctr, nb, &gen);
    double nn, dummy;
    total2(nn, dummy);
    const double inv = 1.0 / sqrt(nn);
    for (int i = gtid; i < V; i += gstride) v[i] *= inv;
    grid_barrier(ctr, nb, &gen);
  }
  double lam = 0;
  int it = 0;
  for (; it < max_iter; ++it) {
    double nn = 0, vw = 0;
    for (int i = gtid; i < V; i += gstride) {
      double acc = 0;
      for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) acc += val[p] * __ldcg(v + col[p]);
      wv[i] = acc;
      nn += acc * acc;
      vw += acc * __ldcg(v + i);
    }
    block_sum2(nn, vw);
    grid_barrier(ctr, nb, &gen);
    total2(nn, vw);
    const double norm = sqrt(nn);
    lam = vw;  // Rayleigh quotient v^T L v (|v| = 1)
    const double old = lam_hist[it & 7];
    __syncthreads();
    if (t == 0) lam_hist[it & 7] = lam;
    const bool stop = norm == 0.0 || (it >= 8 && fabs(lam - old) <= tol * fmax(fabs(lam), 1.0));
    if (!stop) {
      const double inv = 1.0 / norm;
      for (int i = gtid; i < V; i += gstride) v[i] = wv[i] * inv;
    }
    grid_barrier(ctr, nb, &gen);  // (also keeps `partial` from being overwritten while a slow block still sums it)
    if (stop) break;
  }
  if (gtid == 0) out[0] = lam, out[1] = (double)it;
}

// COO output: fp64 Laplacian -> fp32, optionally rescaled to 2 L / lmax - I (fp32 arithmetic, like the reference).
__global__ void emit_coo_kernel(int32_t V, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ val,
                                float scale, int32_t rescale, int64_t* __restrict__ coo_row, int64_t* __restrict__ coo_col,
                                float* __restrict__ coo_val) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
    float x = (float)val[p];
    if (rescale) {
      x = x * scale;
      if (col[p] == i) x -= 1.0f;
    }
    coo_row[p] = i, coo_col[p] = col[p], coo_val[p] = x;
  }
}

__global__ void nested_pool_kernel(int32_t n_coarse, int32_t kernel, int64_t* prow, int64_t* pcol, float* pval, int64_t* urow, int64_t* ucol,
                                   float* uval) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)n_coarse * kernel) return;
  const int64_t c = e / kernel;
  prow[e] = c, pcol[e] = e, pval[e] = 1.0f / (float)kernel;  // pool: row c averages its `kernel` children
  urow[e] = e, ucol[e] = c, uval[e] = 1.0f;                  // unpool: every child copies its parent
}

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

struct GraphWs {
  int32_t* nbr;
  double* nd2;
  int32_t* extra;
  int32_t* cursor;
  int64_t* rowptr;
  int32_t* col;
  double* w;
  double* dinv;
  double* vec0;
  double* vec1;
  double* partial;
  double* scal;  // [0] sigma, [1] lmax (Rayleigh), [2] iterations
  uint32_t *cell, *cell_sorted;
  int32_t *ident, *sorted_pt, *cell_start, *cell_end, *redo, *n_redo;
  void* cub_tmp;
  size_t cub_bytes;
  int32_t G;       // grid cells per axis
  double cell_edge;
  size_t bytes;
};
// Search radius: k + 1 points at density V / (4 pi) cover a cap of chord radius ~ sqrt(4 (k + 1) / V); 1.6x of that
// settles all but a handful of queries inside the 27 cells (the rest take the exact search).
void grid_geometry(int32_t V, int32_t k, int32_t* G, double* edge) {
  const double r = 1.6 * std::sqrt(4.0 * (k + 1.0) / (double)V);
  int32_t g = (int32_t)std::floor(2.0 / std::max(r, 1e-6));
  g = std::max(1, std::min(g, 400));
  *G = g;
  *edge = 2.0 / g;
}
GraphWs carve(void* base, int32_t V, int32_t k) {
  GraphWs g;
  const int64_t cap = (int64_t)V * (2 * k + 1);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? static_cast<char*>(base) + off : nullptr;
    off += align256(bytes);
    return p;
  };
  g.nbr = reinterpret_cast<int32_t*>(take((size_t)V * k * 4));
  g.nd2 = reinterpret_cast<double*>(take((size_t)V * k * 8));
  g.extra = reinterpret_cast<int32_t*>(take((size_t)V * 4));
  g.cursor = reinterpret_cast<int32_t*>(take((size_t)V * 4));
  g.rowptr = reinterpret_cast<int64_t*>(take(((size_t)V + 1) * 8));
  g.col = reinterpret_cast<int32_t*>(take((size_t)cap * 4));
  g.w = reinterpret_cast<double*>(take((size_t)cap * 8));
  g.dinv = reinterpret_cast<double*>(take((size_t)V * 8));
  g.vec0 = reinterpret_cast<double*>(take((size_t)V * 8));
  g.vec1 = reinterpret_cast<double*>(take((size_t)V * 8));
  g.partial = reinterpret_cast<double*>(take(1024 * 8));
  g.scal = reinterpret_cast<double*>(take(64));
  grid_geometry(V, k, &g.G, &g.cell_edge);
  const size_t n_cells = (size_t)g.G * g.G * g.G;
  g.cell = reinterpret_cast<uint32_t*>(take((size_t)V * 4));
  g.cell_sorted = reinterpret_cast<uint32_t*>(take((size_t)V * 4));
  g.ident = reinterpret_cast<int32_t*>(take((size_t)V * 4));
  g.sorted_pt = reinterpret_cast<int32_t*>(take((size_t)V * 4));
  g.cell_start = reinterpret_cast<int32_t*>(take(n_cells * 4));
  g.cell_end = reinterpret_cast<int32_t*>(take(n_cells * 4));
  g.redo = reinterpret_cast<int32_t*>(take((size_t)V * 4));
  g.n_redo = reinterpret_cast<int32_t*>(take(64));
  g.cub_bytes = 0;
  if (cub::DeviceRadixSort::SortPairs(nullptr, g.cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                      (int32_t*)nullptr, V) != cudaSuccess) {
    (void)cudaGetLastError();
    g.cub_bytes = 0;
  }
  g.cub_bytes = std::max(g.cub_bytes, (size_t)V * 16 + ((size_t)1 << 20));  // never less than two key/value double buffers
  g.cub_tmp = take(g.cub_bytes + 256);
  g.bytes = off;
  return g;
}

}  // namespace
}  // namespace dsw

using namespace dsw;

extern "C" {

size_t dsw_graph_workspace_bytes(int32_t V, int32_t k) {
  if (V <= 1 || k < 1) return 0;
  return carve(nullptr, V, k).bytes;
}

int64_t dsw_graph_nnz_capacity(int32_t V, int32_t k) { return (V <= 1 || k < 1) ? 0 : (int64_t)V * (2 * (int64_t)k + 1); }

int dsw_graph_knn_laplacian(const double* xyz, int32_t V, int32_t k, int32_t rescale, double lmax_in, int64_t cap, int64_t* coo_row,
                            int64_t* coo_col, float* coo_val, int64_t* nnz_out, double* lmax_out, void* workspace, size_t workspace_bytes,
                            void* stream) {
  if (!xyz || V <= 1 || k < 1 || !coo_row || !coo_col || !coo_val || !nnz_out) return DSW_ERR_BAD_ARGUMENT;
  if (k >= V || k > KNN_MAX_K) return DSW_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < dsw_graph_workspace_bytes(V, k)) return DSW_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const GraphWs g = carve(workspace, V, k);
  const int64_t E = (int64_t)V * k;
  const int eb = (int)ceil_div64(E, 256), vb = ceil_div(V, 128);

  {
    // k-NN: bin the points into a uniform grid, search the 27 cells around each query, exact search for the leftovers
    const size_t n_cells = (size_t)g.G * g.G * g.G;
    const double inv_cell = 1.0 / g.cell_edge;
    cell_id_kernel<<<vb, 128, 0, st>>>(xyz, V, inv_cell, g.G, g.cell, g.ident);
    size_t tmp = g.cub_bytes;
    int bits = 1;
    while (((size_t)1 << bits) < n_cells) ++bits;
    if (cub::DeviceRadixSort::SortPairs(g.cub_tmp, tmp, g.cell, g.cell_sorted, g.ident, g.sorted_pt, V, 0, bits, st) != cudaSuccess)
      return DSW_ERR_CUDA;
    DSW_CUDA_TRY(cudaMemsetAsync(g.cell_start, 0, n_cells * 4, st));
    DSW_CUDA_TRY(cudaMemsetAsync(g.cell_end, 0, n_cells * 4, st));
    DSW_CUDA_TRY(cudaMemsetAsync(g.n_redo, 0, 4, st));
    cell_start_kernel<<<vb, 128, 0, st>>>(g.cell_sorted, V, g.cell_start, g.cell_end);
    knn_grid_kernel<<<ceil_div(V, KNN_THREADS), KNN_THREADS, 0, st>>>(xyz, g.sorted_pt, V, k, inv_cell, g.G, g.cell_edge * g.cell_edge,
                                                                    g.cell_start, g.cell_end, g.nbr, g.nd2, g.redo, g.n_redo);
    knn_brute_kernel<<<ceil_div(V, KNN_THREADS), KNN_THREADS, 0, st>>>(xyz, V, k, g.redo, g.n_redo, g.nbr, g.nd2);
    DSW_TRY(check_launch());
  }
  const int nparts = (int)std::min<int64_t>(1024, ceil_div64(E, 256));
  dist_partial_kernel<<<nparts, 256, 0, st>>>(g.nd2, E, g.partial);
  sigma_kernel<<<1, 1, 0, st>>>(g.partial, nparts, E, g.scal);
  DSW_CUDA_TRY(cudaMemsetAsync(g.extra, 0, (size_t)V * 4, st));
  DSW_CUDA_TRY(cudaMemsetAsync(g.cursor, 0, (size_t)V * 4, st));
  count_missing_kernel<<<eb, 256, 0, st>>>(g.nbr, V, k, g.extra);
  rowptr_kernel<<<1, 1024, 0, st>>>(g.extra, V, k, g.rowptr);
  fill_kernel<<<eb, 256, 0, st>>>(g.nbr, g.nd2, V, k, g.rowptr, g.cursor, g.col, g.w);
  sort_weight_kernel<<<vb, 128, 0, st>>>(V, g.rowptr, g.col, g.w, g.scal, g.dinv);
  laplacian_kernel<<<vb, 128, 0, st>>>(V, g.rowptr, g.col, g.w, g.dinv);
  DSW_TRY(check_launch());

  int64_t nnz = 0;
  DSW_CUDA_TRY(cudaMemcpyAsync(&nnz, g.rowptr + V, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  DSW_CUDA_TRY(cudaStreamSynchronize(st));
  if (nnz > cap) return DSW_ERR_WORKSPACE;
  *nnz_out = nnz;

  double lmax = lmax_in;
  if (rescale && !(lmax > 0.0)) {
    int dev = 0, n_sm = 1;
    DSW_CUDA_TRY(cudaGetDevice(&dev));
    DSW_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const int nb = std::max(1, std::min(n_sm, ceil_div(V, PI_THREADS)));  // one block per SM: all co-resident (grid barrier)
    DSW_CUDA_TRY(cudaMemsetAsync(g.n_redo, 0, 64, st));                  // the barrier counter (the k-NN pass is done with it)
    power_iteration_kernel<<<nb, PI_THREADS, 0, st>>>(V, g.rowptr, g.col, g.w, g.vec0, g.vec1, g.partial,
                                                      reinterpret_cast<unsigned int*>(g.n_redo), 50000, 1e-10, g.scal + 1);
    DSW_TRY(check_launch());
    double res[2] = {0, 0};
    DSW_CUDA_TRY(cudaMemcpyAsync(res, g.scal + 1, sizeof(res), cudaMemcpyDeviceToHost, st));
    DSW_CUDA_TRY(cudaStreamSynchronize(st));
    lmax = res[0] * (1.0 + 2.0 * 5e-3);  // the reference's safety margin (layers.py:66-68)
  }
  if (lmax_out) *lmax_out = lmax;
  emit_coo_kernel<<<vb, 128, 0, st>>>(V, g.rowptr, g.col, g.w, rescale ? (float)(2.0 / lmax) : 1.0f, rescale, coo_row, coo_col, coo_val);
  return check_launch();
}

int dsw_graph_nested_pool(int32_t n_fine, int32_t kernel, int64_t* pool_row, int64_t* pool_col, float* pool_val, int64_t* unpool_row,
                          int64_t* unpool_col, float* unpool_val, void* stream) {
  if (n_fine <= 0 || kernel <= 0 || n_fine % kernel || !pool_row || !pool_col || !pool_val || !unpool_row || !unpool_col || !unpool_val)
    return DSW_ERR_BAD_ARGUMENT;
  const int32_t n_coarse = n_fine / kernel;
  nested_pool_kernel<<<ceil_div(n_fine, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n_coarse, kernel, pool_row, pool_col, pool_val,
                                                                                          unpool_row, unpool_col, unpool_val);
  return check_launch();
}

}  // extern "C"
