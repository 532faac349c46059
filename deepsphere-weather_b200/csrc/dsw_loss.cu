// Area-weighted MSE loss of the reference's training loop (modules/loss.py:118-148 WeightedMSELoss.forward):
//   weighted[b][v][f] = w[v] * (pred - label)^2;  "mean": sum / sum(w) / B / F;  "sum": sum (the reference multiplies by
//   len(weights) after reshaping the weights to [1, V, 1], i.e. by 1: loss.py:141-144);  "none": weighted
// and its gradient.  One streaming pass each way (HBM-bound, tiny next to the network); the scalar reductions
// are fixed-order two-stage sums (deterministic).
#include <algorithm>

#include "dsw_internal.cuh"

namespace dsw {

constexpr int LS_THREADS = 256;
constexpr int LS_BLOCKS = 148 * 2;

__global__ void __launch_bounds__(LS_THREADS) wmse_partial_kernel(const float* __restrict__ p, const float* __restrict__ l,
                                                                  const float* __restrict__ w, int64_t n, int32_t V, int32_t F,
                                                                  float* __restrict__ partial) {
  __shared__ float red[LS_THREADS / 32];
  pdl_trigger();
  pdl_wait();
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = p[i] - l[i];
    const float ww = w ? __ldg(w + (i / F) % V) : 1.f;
    acc = fmaf(ww * d, d, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LS_THREADS / 32; ++i) s += red[i];
    partial[blockIdx.x] = s;
  }
}

// out2[0] = loss, out2[1] = d loss / d (weighted sum of squares)
__global__ void __launch_bounds__(LS_THREADS) wmse_final_kernel(const float* __restrict__ partial, int32_t np,
                                                                const float* __restrict__ w, int32_t B, int32_t V, int32_t F,
                                                                int32_t reduction, float* __restrict__ loss, float* __restrict__ out2) {
  __shared__ double red[2][LS_THREADS];
  pdl_trigger();
  pdl_wait();
  double s = 0.0, sw = 0.0;
  for (int i = threadIdx.x; i < np; i += LS_THREADS) s += (double)partial[i];
  if (w)
    for (int i = threadIdx.x; i < V; i += LS_THREADS) sw += (double)__ldg(w + i);
  red[0][threadIdx.x] = s, red[1][threadIdx.x] = sw;
  __syncthreads();
  for (int o = LS_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[0][threadIdx.x] += red[0][threadIdx.x + o], red[1][threadIdx.x] += red[1][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double sum_w = w ? red[1][0] : (double)V;
    const double scale = reduction == 1 ? 1.0 : 1.0 / sum_w / (double)B / (double)F;
    loss[0] = (float)(red[0][0] * scale);
    out2[0] = loss[0];
    out2[1] = (float)scale;
  }
}

// grad[i] = g * 2 * w[v] * (pred - label);  g = gout[0] * scale[0] (scalar losses) or gout[i] (reduction "none")
__global__ void __launch_bounds__(LS_THREADS) wmse_bwd_kernel(const float* __restrict__ p, const float* __restrict__ l,
                                                              const float* __restrict__ w, const float* __restrict__ gout,
                                                              const float* __restrict__ scale, int64_t n, int32_t V, int32_t F,
                                                              float* __restrict__ grad) {
  pdl_trigger();
  pdl_wait();
  const float gs = scale ? __ldg(gout) * __ldg(scale) : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float ww = w ? __ldg(w + (i / F) % V) : 1.f;
    const float g = scale ? gs : gout[i];
    grad[i] = 2.f * g * ww * (p[i] - l[i]);
  }
}

__global__ void __launch_bounds__(LS_THREADS) wmse_none_kernel(const float* __restrict__ p, const float* __restrict__ l,
                                                               const float* __restrict__ w, int64_t n, int32_t V, int32_t F,
                                                               float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = p[i] - l[i];
    out[i] = (w ? __ldg(w + (i / F) % V) : 1.f) * d * d;
  }
}

static int ls_blocks(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(LS_BLOCKS, ceil_div64(n, LS_THREADS))); }

}  // namespace dsw

using namespace dsw;

extern "C" {

size_t dsw_wmse_workspace_bytes(void) { return (size_t)(LS_BLOCKS + 2) * sizeof(float); }

int dsw_wmse_fwd(const float* pred, const float* label, const float* weights, float* loss, void* workspace, size_t workspace_bytes,
                 int32_t B, int32_t V, int32_t F, int32_t reduction, void* stream) {
  if (!pred || !label || !loss || B <= 0 || V <= 0 || F <= 0 || reduction < 0 || reduction > 1) return DSW_ERR_BAD_ARGUMENT;
  if (!workspace || workspace_bytes < dsw_wmse_workspace_bytes()) return DSW_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n = (int64_t)B * V * F;
  float* ws = static_cast<float*>(workspace);
  const int nb = ls_blocks(n);
  DSW_CUDA_TRY(launch_pdl(wmse_partial_kernel, dim3(nb), dim3(LS_THREADS), 0, st, pdl_enabled(), pred, label, weights, n, V, F, ws + 2));
  DSW_TRY(check_launch());
  DSW_CUDA_TRY(launch_pdl(wmse_final_kernel, dim3(1), dim3(LS_THREADS), 0, st, pdl_enabled(), (const float*)(ws + 2), (int32_t)nb, weights, B,
                          V, F, reduction, loss, ws));
  return check_launch();
}

int dsw_wmse_bwd(const float* pred, const float* label, const float* weights, const void* workspace, const float* grad_out,
                 float* grad_pred, int32_t B, int32_t V, int32_t F, void* stream) {
  if (!pred || !label || !workspace || !grad_out || !grad_pred || B <= 0 || V <= 0 || F <= 0) return DSW_ERR_BAD_ARGUMENT;
  const int64_t n = (int64_t)B * V * F;
  const float* ws = static_cast<const float*>(workspace);
  DSW_CUDA_TRY(launch_pdl(wmse_bwd_kernel, dim3(ls_blocks(n)), dim3(LS_THREADS), 0, static_cast<cudaStream_t>(stream), pdl_enabled(), pred,
                          label, weights, grad_out, ws + 1, n, V, F, grad_pred));
  return check_launch();
}

int dsw_wmse_none_fwd(const float* pred, const float* label, const float* weights, float* out, int32_t B, int32_t V, int32_t F,
                      void* stream) {
  if (!pred || !label || !out || B <= 0 || V <= 0 || F <= 0) return DSW_ERR_BAD_ARGUMENT;
  const int64_t n = (int64_t)B * V * F;
  DSW_CUDA_TRY(launch_pdl(wmse_none_kernel, dim3(ls_blocks(n)), dim3(LS_THREADS), 0, static_cast<cudaStream_t>(stream), pdl_enabled(), pred,
                          label, weights, n, V, F, out));
  return check_launch();
}

int dsw_wmse_none_bwd(const float* pred, const float* label, const float* weights, const float* grad_out, float* grad_pred, int32_t B,
                      int32_t V, int32_t F, void* stream) {
  if (!pred || !label || !grad_out || !grad_pred || B <= 0 || V <= 0 || F <= 0) return DSW_ERR_BAD_ARGUMENT;
  const int64_t n = (int64_t)B * V * F;
  DSW_CUDA_TRY(launch_pdl(wmse_bwd_kernel, dim3(ls_blocks(n)), dim3(LS_THREADS), 0, static_cast<cudaStream_t>(stream), pdl_enabled(), pred,
                          label, weights, grad_out, (const float*)nullptr, n, V, F, grad_pred));
  return check_launch();
}

}  // extern "C"
