// Index-producing / index-consuming pools (bit-exact integer + gather/scatter work).
//
//  * max-value pooling over a sparse remap matrix  — reference GeneralMaxValPool/Unpool,
//    modules/layers.py:1040-1103
//  * nested-order (HEALPix) max / avg pools        — reference modules/layers.py:784-941
//
// All kernels work on 32(node) x 32(feature) tiles of one sample.  Feature-fastest threads keep the
// channel-last value accesses coalesced; the int64 index planes are node-fastest in memory
// ([F*B][Vc] and [B][F][Vc]), so indices pass through a padded shared-memory tile and are read /
// written node-fastest.  HBM-bound; no data reuse to exploit beyond that.
#include "dsw_internal.cuh"

namespace dsw {

namespace {
constexpr int TILE = 32;
constexpr int TROWS = 8;  // blockDim = (32, 8); each thread covers TILE/TROWS rows of the tile
}

// torch.argmax / max_pool1d semantics: first maximum wins; NaN beats everything (first NaN wins).
__device__ __forceinline__ bool better(float cand, float best) {
  return (cand > best) || (cand != cand && best == best);
}

__global__ void __launch_bounds__(256) maxval_pool_fwd_kernel(const int32_t* __restrict__ rowptr,
                                                              const int32_t* __restrict__ col,
                                                              const float* __restrict__ val, int32_t Vc,
                                                              const float* __restrict__ x, int64_t x_sB,
                                                              int64_t x_sV, float* __restrict__ y,
                                                              int64_t* __restrict__ idx_row,
                                                              int64_t* __restrict__ idx_col, int32_t B,
                                                              int32_t F) {
  __shared__ int32_t pick[TILE][TILE + 1];  // [node][feature]
  const int b = blockIdx.z;
  const int r0 = blockIdx.x * TILE, f0 = blockIdx.y * TILE;
  const int f = f0 + threadIdx.x;
  for (int i = threadIdx.y; i < TILE; i += TROWS) {
    const int r = r0 + i;
    int32_t best_j = 0;
    if (r < Vc && f < F) {
      const int e0 = __ldg(rowptr + r), e1 = __ldg(rowptr + r + 1);
      float best = 0.f;
      for (int e = e0; e < e1; ++e) {
        const int j = __ldg(col + e);
        const float cand = __ldg(val + e) * __ldg(x + b * x_sB + j * x_sV + f);
        if (e == e0 || better(cand, best)) best = cand, best_j = j;
      }
      y[((int64_t)b * Vc + r) * F + f] = __ldg(x + b * x_sB + best_j * x_sV + f);
    }
    pick[i][threadIdx.x] = best_j;
  }
  __syncthreads();
  // node-fastest write of the index planes: i = (f*B + b)*Vc + r
  const int r = r0 + threadIdx.x;
  for (int i = threadIdx.y; i < TILE; i += TROWS) {
    const int ff = f0 + i;
    if (r < Vc && ff < F) {
      const int64_t c = (int64_t)ff * B + b;
      idx_row[c * Vc + r] = pick[threadIdx.x][i];
      idx_col[c * Vc + r] = c;
    }
  }
}

// MODE 0: dst[b, idx_row[i], f] (+)= src[b, r, f]       (scatter; ATOMIC selects accumulate)
// MODE 1: dst[b, r, f]          = src[b, idx_row[i], f]  (gather)
// with i = (f*B + b)*Vc + r — the reference's nnz_ind layout (layers.py:1075-1079).  idx_col is
// implied by i (column c = f*B + b) and is validated in debug builds only.
template <int MODE, bool ATOMIC>
__global__ void __launch_bounds__(256) maxval_index_kernel(const float* __restrict__ src,
                                                           const int64_t* __restrict__ idx_row,
                                                           float* __restrict__ dst, int32_t B, int32_t V,
                                                           int32_t Vc, int32_t F) {
  __shared__ int32_t pick[TILE][TILE + 1];
  const int b = blockIdx.z;
  const int r0 = blockIdx.x * TILE, f0 = blockIdx.y * TILE;
  {
    const int r = r0 + threadIdx.x;
    for (int i = threadIdx.y; i < TILE; i += TROWS) {
      const int ff = f0 + i;
      int32_t j = 0;
      if (r < Vc && ff < F) j = (int32_t)__ldg(idx_row + ((int64_t)ff * B + b) * Vc + r);
      pick[threadIdx.x][i] = j;
    }
  }
  __syncthreads();
  const int f = f0 + threadIdx.x;
  if (f >= F) return;
  for (int i = threadIdx.y; i < TILE; i += TROWS) {
    const int r = r0 + i;
    if (r >= Vc) break;
    const int64_t fine = ((int64_t)b * V + pick[i][threadIdx.x]) * F + f;
    const int64_t coarse = ((int64_t)b * Vc + r) * F + f;
    if (MODE == 0) {
      if (ATOMIC)
        atomicAdd(dst + fine, __ldg(src + coarse));
      else
        dst[fine] = __ldg(src + coarse);
    } else {
      dst[coarse] = __ldg(src + fine);
    }
  }
}

// ---- nested-order pools ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nested_maxpool_kernel(const float* __restrict__ x, int64_t x_sB,
                                                             int64_t x_sV, float* __restrict__ y,
                                                             int64_t* __restrict__ idx, int32_t B,
                                                             int32_t Vc, int32_t F, int32_t kernel) {
  __shared__ int32_t pick[TILE][TILE + 1];
  const int b = blockIdx.z;
  const int r0 = blockIdx.x * TILE, f0 = blockIdx.y * TILE;
  const int f = f0 + threadIdx.x;
  for (int i = threadIdx.y; i < TILE; i += TROWS) {
    const int r = r0 + i;
    int32_t best_j = 0;
    if (r < Vc && f < F) {
      const float* __restrict__ p = x + b * x_sB + (int64_t)r * kernel * x_sV + f;
      float best = __ldg(p);
      best_j = r * kernel;
      for (int q = 1; q < kernel; ++q) {
        const float cand = __ldg(p + q * x_sV);
        if (better(cand, best)) best = cand, best_j = r * kernel + q;
      }
      y[((int64_t)b * Vc + r) * F + f] = best;
    }
    pick[i][threadIdx.x] = best_j;
  }
  __syncthreads();
  const int r = r0 + threadIdx.x;
  for (int i = threadIdx.y; i < TILE; i += TROWS) {
    const int ff = f0 + i;
    if (r < Vc && ff < F) idx[((int64_t)b * F + ff) * Vc + r] = pick[threadIdx.x][i];
  }
}

// MODE 0: dst[b, idx[b,f,r], f] = src[b,r,f] (dst pre-zeroed);  MODE 1: dst[b,r,f] = src[b, idx[b,f,r], f]
template <int MODE>
__global__ void __launch_bounds__(256) nested_index_kernel(const float* __restrict__ src,
                                                           const int64_t* __restrict__ idx,
                                                           float* __restrict__ dst, int32_t B, int32_t V,
                                                           int32_t Vc, int32_t F) {
  __shared__ int32_t pick[TILE][TILE + 1];
  const int b = blockIdx.z;
  const int r0 = blockIdx.x * TILE, f0 = blockIdx.y * TILE;
  {
    const int r = r0 + threadIdx.x;
    for (int i = threadIdx.y; i < TILE; i += TROWS) {
      const int ff = f0 + i;
      int32_t j = 0;
      if (r < Vc && ff < F) j = (int32_t)__ldg(idx + ((int64_t)b * F + ff) * Vc + r);
      pick[threadIdx.x][i] = j;
    }
  }
  __syncthreads();
  const int f = f0 + threadIdx.x;
  if (f >= F) return;
  for (int i = threadIdx.y; i < TILE; i += TROWS) {
    const int r = r0 + i;
    if (r >= Vc) break;
    const int64_t fine = ((int64_t)b * V + pick[i][threadIdx.x]) * F + f;
    const int64_t coarse = ((int64_t)b * Vc + r) * F + f;
    if (MODE == 0)
      dst[fine] = __ldg(src + coarse);
    else
      dst[coarse] = __ldg(src + fine);
  }
}

// y[b,r,f] = (sum_{q<kernel} x[b, r*kernel+q, f]) * scale, summed left to right
__global__ void __launch_bounds__(256) nested_sum_kernel(const float* __restrict__ x, int64_t x_sB,
                                                         int64_t x_sV, float* __restrict__ y, int32_t Vc,
                                                         int32_t F, int32_t kernel, float divisor) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= (int64_t)Vc * F) return;
  const int r = (int)(i / F), f = (int)(i - (int64_t)r * F);
  const float* __restrict__ p = x + b * x_sB + (int64_t)r * kernel * x_sV + f;
  float s = __ldg(p);
  for (int q = 1; q < kernel; ++q) s += __ldg(p + q * x_sV);
  y[(int64_t)b * Vc * F + i] = (divisor == 1.f) ? s : s / divisor;
}

// y[b,v,f] = scale * x[b, v / kernel, f]
__global__ void __launch_bounds__(256) nested_repeat_kernel(const float* __restrict__ x, int64_t x_sB,
                                                            int64_t x_sV, float* __restrict__ y, int32_t V,
                                                            int32_t F, int32_t kernel, float scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= (int64_t)V * F) return;
  const int v = (int)(i / F), f = (int)(i - (int64_t)v * F);
  const float s = __ldg(x + b * x_sB + (int64_t)(v / kernel) * x_sV + f);
  y[(int64_t)b * V * F + i] = (scale == 1.f) ? s : s * scale;
}

static dim3 tile_grid(int32_t Vc, int32_t F, int32_t B) { return dim3(ceil_div(Vc, TILE), ceil_div(F, TILE), B); }

}  // namespace dsw

using namespace dsw;

extern "C" {

int dsw_maxval_pool_fwd(const dsw_plan* mat, const float* x, int64_t x_sB, int64_t x_sV, float* y,
                        int64_t* idx_row, int64_t* idx_col, int32_t B, int32_t F, void* stream) {
  if (!mat || !x || !y || !idx_row || !idx_col || B <= 0 || F <= 0) return DSW_ERR_BAD_ARGUMENT;
  if (B > 65535) return DSW_ERR_UNSUPPORTED;
  const dsw_csr& m = mat->fwd;
  maxval_pool_fwd_kernel<<<tile_grid(m.n_rows, F, B), dim3(TILE, TROWS), 0, static_cast<cudaStream_t>(stream)>>>(
      m.rowptr, m.col, m.val, m.n_rows, x, x_sB, x_sV, y, idx_row, idx_col, B, F);
  return check_launch();
}

int dsw_maxval_pool_bwd(const float* dy, const int64_t* idx_row, float* dx, int32_t B, int32_t V, int32_t Vc,
                        int32_t F, void* stream) {
  if (!dy || !idx_row || !dx || B <= 0 || V <= 0 || Vc <= 0 || F <= 0) return DSW_ERR_BAD_ARGUMENT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DSW_CUDA_TRY(cudaMemsetAsync(dx, 0, (size_t)B * V * F * sizeof(float), st));
  maxval_index_kernel<0, true><<<tile_grid(Vc, F, B), dim3(TILE, TROWS), 0, st>>>(dy, idx_row, dx, B, V, Vc, F);
  return check_launch();
}

int dsw_scatter_unpool_fwd(const float* x, const int64_t* idx_row, const int64_t* idx_col, float* out,
                           int32_t B, int32_t V, int32_t Vc, int32_t F, void* stream) {
  (void)idx_col;  // column index is implied by position: idx_col[i] == i / Vc (layers.py:1075-1079)
  if (!x || !idx_row || !out || B <= 0 || V <= 0 || Vc <= 0 || F <= 0) return DSW_ERR_BAD_ARGUMENT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DSW_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)B * V * F * sizeof(float), st));
  maxval_index_kernel<0, false><<<tile_grid(Vc, F, B), dim3(TILE, TROWS), 0, st>>>(x, idx_row, out, B, V, Vc, F);
  return check_launch();
}

int dsw_scatter_unpool_bwd(const float* dout, const int64_t* idx_row, const int64_t* idx_col, float* dx,
                           int32_t B, int32_t V, int32_t Vc, int32_t F, void* stream) {
  (void)idx_col;
  if (!dout || !idx_row || !dx || B <= 0 || V <= 0 || Vc <= 0 || F <= 0) return DSW_ERR_BAD_ARGUMENT;
  maxval_index_kernel<1, false><<<tile_grid(Vc, F, B), dim3(TILE, TROWS), 0, static_cast<cudaStream_t>(stream)>>>(
      dout, idx_row, dx, B, V, Vc, F);
  return check_launch();
}

int dsw_nested_maxpool_fwd(const float* x, int64_t x_sB, int64_t x_sV, float* y, int64_t* idx, int32_t B,
                           int32_t V, int32_t F, int32_t kernel, void* stream) {
  if (!x || !y || !idx || B <= 0 || V <= 0 || F <= 0 || kernel <= 0) return DSW_ERR_BAD_ARGUMENT;
  if (V % kernel) return DSW_ERR_SHAPE;
  const int32_t Vc = V / kernel;
  nested_maxpool_kernel<<<tile_grid(Vc, F, B), dim3(TILE, TROWS), 0, static_cast<cudaStream_t>(stream)>>>(
      x, x_sB, x_sV, y, idx, B, Vc, F, kernel);
  return check_launch();
}

int dsw_nested_scatter(const float* src, const int64_t* idx, float* dst, int32_t B, int32_t V, int32_t F,
                       int32_t kernel, void* stream) {
  if (!src || !idx || !dst || B <= 0 || V <= 0 || F <= 0 || kernel <= 0) return DSW_ERR_BAD_ARGUMENT;
  if (V % kernel) return DSW_ERR_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DSW_CUDA_TRY(cudaMemsetAsync(dst, 0, (size_t)B * V * F * sizeof(float), st));
  nested_index_kernel<0><<<tile_grid(V / kernel, F, B), dim3(TILE, TROWS), 0, st>>>(src, idx, dst, B, V, V / kernel, F);
  return check_launch();
}

int dsw_nested_gather(const float* src, const int64_t* idx, float* dst, int32_t B, int32_t V, int32_t F,
                      int32_t kernel, void* stream) {
  if (!src || !idx || !dst || B <= 0 || V <= 0 || F <= 0 || kernel <= 0) return DSW_ERR_BAD_ARGUMENT;
  if (V % kernel) return DSW_ERR_SHAPE;
  nested_index_kernel<1><<<tile_grid(V / kernel, F, B), dim3(TILE, TROWS), 0, static_cast<cudaStream_t>(stream)>>>(
      src, idx, dst, B, V, V / kernel, F);
  return check_launch();
}

int dsw_nested_avgpool_fwd(const float* x, int64_t x_sB, int64_t x_sV, float* y, int32_t B, int32_t V,
                           int32_t F, int32_t kernel, void* stream) {
  if (!x || !y || B <= 0 || V <= 0 || F <= 0 || kernel <= 0) return DSW_ERR_BAD_ARGUMENT;
  if (V % kernel) return DSW_ERR_SHAPE;
  const int32_t Vc = V / kernel;
  dim3 grid((unsigned)ceil_div64((int64_t)Vc * F, 256), B);
  nested_sum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_sB, x_sV, y, Vc, F, kernel,
                                                                        (float)kernel);
  return check_launch();
}

int dsw_nested_sum(const float* x, int64_t x_sB, int64_t x_sV, float* y, int32_t B, int32_t V, int32_t F,
                   int32_t kernel, void* stream) {
  if (!x || !y || B <= 0 || V <= 0 || F <= 0 || kernel <= 0) return DSW_ERR_BAD_ARGUMENT;
  if (V % kernel) return DSW_ERR_SHAPE;
  const int32_t Vc = V / kernel;
  dim3 grid((unsigned)ceil_div64((int64_t)Vc * F, 256), B);
  nested_sum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_sB, x_sV, y, Vc, F, kernel, 1.f);
  return check_launch();
}

int dsw_nested_repeat(const float* x, int64_t x_sB, int64_t x_sV, float* y, float scale, int32_t B, int32_t V,
                      int32_t F, int32_t kernel, void* stream) {
  if (!x || !y || B <= 0 || V <= 0 || F <= 0 || kernel <= 0) return DSW_ERR_BAD_ARGUMENT;
  if (V % kernel) return DSW_ERR_SHAPE;
  dim3 grid((unsigned)ceil_div64((int64_t)V * F, 256), B);
  nested_repeat_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_sB, x_sV, y, V, F, kernel, scale);
  return check_launch();
}

}  // extern "C"
