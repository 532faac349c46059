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

#include "dsw_internal.cuh"

namespace dsw {
namespace {

constexpr int KNN_MAX_K = 64;
constexpr int KNN_THREADS = 128;

// One thread per query point; candidates stream through shared memory in tiles.  The k best (d2, idx) pairs are kept
// sorted ascending in local arrays (insertions are rare once the list has warmed up).
__global__ void __launch_bounds__(KNN_THREADS) knn_kernel(const double* __restrict__ xyz, int32_t V, int32_t k, int32_t* __restrict__ nbr,
                                                           double* __restrict__ nd2) {
  __shared__ double sx[KNN_THREADS], sy[KNN_THREADS], sz[KNN_THREADS];
  const int q = blockIdx.x * KNN_THREADS + threadIdx.x;
  const bool live = q < V;
  double qx = 0, qy = 0, qz = 0;
  if (live) qx = xyz[3 * (size_t)q], qy = xyz[3 * (size_t)q + 1], qz = xyz[3 * (size_t)q + 2];
  double bd[KNN_MAX_K];
  int32_t bi[KNN_MAX_K];
  for (int i = 0; i < k; ++i) bd[i] = 1e300, bi[i] = 0x7fffffff;
  double worst = 1e300;  // bd[k - 1], kept in a register
  for (int base = 0; base < V; base += KNN_THREADS) {
    const int c = base + threadIdx.x;
    if (c < V) sx[threadIdx.x] = xyz[3 * (size_t)c], sy[threadIdx.x] = xyz[3 * (size_t)c + 1], sz[threadIdx.x] = xyz[3 * (size_t)c + 2];
    __syncthreads();
    const int n = min(KNN_THREADS, V - base);
    if (live) {
      for (int j = 0; j < n; ++j) {
        const int cj = base + j;
        if (cj == q) continue;
        const double dx = qx - sx[j], dy = qy - sy[j], dz = qz - sz[j];
        // no fused multiply-add: the squared chord length is rounded exactly like numpy's ((a - b) ** 2).sum(-1), so
        // that ties fall the same way on both sides of a parity test
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        // candidates arrive in ascending index order: a tie with the current worst keeps the earlier (lower) index
        if (d2 < worst) {
          int p = k - 1;
          while (p > 0 && bd[p - 1] > d2) {
            bd[p] = bd[p - 1], bi[p] = bi[p - 1];
            --p;
          }
          bd[p] = d2, bi[p] = cj;
          worst = bd[k - 1];
        }
      }
    }
    __syncthreads();
  }
  if (live)
    for (int i = 0; i < k; ++i) nbr[(size_t)q * k + i] = bi[i], nd2[(size_t)q * k + i] = bd[i];
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

// Largest eigenvalue by power iteration in ONE launch (a single block: the iterations are strictly sequential and the
// operator is tiny — 49 152 x 22 at nside 64).  Fixed start vector, fixed reduction tree: reproducible.  Stops when the
// Rayleigh quotient has moved by less than `tol` (relative) over 8 iterations, or after `max_iter`.
__global__ void __launch_bounds__(1024) power_iteration_kernel(int32_t V, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                               const double* __restrict__ val, double* __restrict__ v,
                                                               double* __restrict__ wv, int32_t max_iter, double tol, double* __restrict__ out) {
  __shared__ double red[2][1024];
  __shared__ double lam_hist[8];
  const int t = threadIdx.x;
  for (int i = t; i < V; i += 1024) v[i] = cos((double)i * 0.7390851332151607) + 1.5;
  __syncthreads();
  {
    double acc = 0;
    for (int i = t; i < V; i += 1024) acc += v[i] * v[i];
    red[0][t] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
      if (t < s) red[0][t] += red[0][t + s];
      __syncthreads();
    }
    const double inv = 1.0 / sqrt(red[0][0]);
    __syncthreads();
    for (int i = t; i < V; i += 1024) v[i] *= inv;
    __syncthreads();
  }
  double lam = 0;
  int it = 0;
  for (; it < max_iter; ++it) {
    double nn = 0, vw = 0;
    for (int i = t; i < V; i += 1024) {
      double acc = 0;
      for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) acc += val[p] * v[col[p]];
      wv[i] = acc;
      nn += acc * acc;
      vw += acc * v[i];
    }
    red[0][t] = nn, red[1][t] = vw;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
      if (t < s) red[0][t] += red[0][t + s], red[1][t] += red[1][t + s];
      __syncthreads();
    }
    const double norm = sqrt(red[0][0]);
    lam = red[1][0];  // Rayleigh quotient v^T L v (|v| = 1)
    const double old = lam_hist[it & 7];
    __syncthreads();
    if (t == 0) lam_hist[it & 7] = lam;
    if (norm == 0.0) break;
    const double inv = 1.0 / norm;
    for (int i = t; i < V; i += 1024) v[i] = wv[i] * inv;
    __syncthreads();
    if (it >= 8 && fabs(lam - old) <= tol * fmax(fabs(lam), 1.0)) break;
  }
  if (t == 0) out[0] = lam, out[1] = (double)it;
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
  size_t bytes;
};
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

  knn_kernel<<<ceil_div(V, KNN_THREADS), KNN_THREADS, 0, st>>>(xyz, V, k, g.nbr, g.nd2);
  DSW_TRY(check_launch());
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
    power_iteration_kernel<<<1, 1024, 0, st>>>(V, g.rowptr, g.col, g.w, g.vec0, g.vec1, 20000, 1e-10, g.scal + 1);
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
