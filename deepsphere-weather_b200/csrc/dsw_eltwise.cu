// ResBlock tail:  y = w * conv_out + skip   (ReZero scale + residual add) and its gradients in one pass
// each.  Replaces the two in-place elementwise passes of the reference
// (modules/my_models_graph.py:211-215: `x_out *= self.rezero_weight; x_out += self.res_connection(x)`)
// and the four that autograd derives from them (clone of x_out, MulBackward's two products, the
// reduction to the scalar).  Pure streaming, HBM-bound: forward 3 planes, backward 3 planes.
//
// The scalar gradient is a fixed-order two-stage reduction (per-CTA partials, then one CTA), so the
// result is deterministic run to run.
#include <algorithm>

#include "dsw_internal.cuh"

namespace dsw {

constexpr int EW_THREADS = 256;
constexpr int EW_BLOCKS = 148 * 4;  // partial sums of the backward reduction

__global__ void __launch_bounds__(EW_THREADS) rezero_fwd_kernel(const float* __restrict__ a, const float* __restrict__ s,
                                                                const float* __restrict__ w, float* __restrict__ y,
                                                                int64_t n4, int64_t n) {
  pdl_trigger();
  pdl_wait();
  const float ww = __ldg(w);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = t; i < n4; i += stride) {
    const float4 av = __ldcs(reinterpret_cast<const float4*>(a) + i);
    const float4 sv = __ldcs(reinterpret_cast<const float4*>(s) + i);
    float4 o;
    o.x = fmaf(av.x, ww, sv.x), o.y = fmaf(av.y, ww, sv.y), o.z = fmaf(av.z, ww, sv.z), o.w = fmaf(av.w, ww, sv.w);
    reinterpret_cast<float4*>(y)[i] = o;
  }
  for (int64_t i = n4 * 4 + t; i < n; i += stride) y[i] = fmaf(a[i], ww, s[i]);
}

__global__ void __launch_bounds__(EW_THREADS) rezero_bwd_kernel(const float* __restrict__ g, const float* __restrict__ a,
                                                                const float* __restrict__ w, float* __restrict__ da,
                                                                float* __restrict__ partial, int64_t n4, int64_t n) {
  __shared__ float red[EW_THREADS / 32];
  pdl_trigger();
  pdl_wait();
  const float ww = __ldg(w);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  for (int64_t i = t; i < n4; i += stride) {
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
    if (partial) {
      const float4 av = __ldcs(reinterpret_cast<const float4*>(a) + i);
      acc = fmaf(gv.x, av.x, acc), acc = fmaf(gv.y, av.y, acc), acc = fmaf(gv.z, av.z, acc), acc = fmaf(gv.w, av.w, acc);
    }
    if (da) reinterpret_cast<float4*>(da)[i] = make_float4(gv.x * ww, gv.y * ww, gv.z * ww, gv.w * ww);
  }
  for (int64_t i = n4 * 4 + t; i < n; i += stride) {
    const float gv = g[i];
    if (partial) acc = fmaf(gv, a[i], acc);
    if (da) da[i] = gv * ww;
  }
  if (!partial) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < EW_THREADS / 32; ++i) s += red[i];
    partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(EW_THREADS) rezero_reduce_kernel(const float* __restrict__ partial, int32_t np,
                                                                   float* __restrict__ dw) {
  __shared__ double red[EW_THREADS];
  pdl_trigger();
  pdl_wait();
  double s = 0.0;
  for (int i = threadIdx.x; i < np; i += EW_THREADS) s += (double)partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = EW_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) dw[0] = (float)red[0];
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace dsw

using namespace dsw;

extern "C" {

size_t dsw_rezero_bwd_workspace_bytes(void) { return (size_t)EW_BLOCKS * sizeof(float); }

int dsw_rezero_fwd(const float* conv_out, const float* skip, const float* w, float* y, int64_t n, void* stream) {
  if (!conv_out || !skip || !w || !y || n <= 0) return DSW_ERR_BAD_ARGUMENT;
  const bool v4 = al16(conv_out) && al16(skip) && al16(y);
  const int64_t n4 = v4 ? n / 4 : 0;
  const int blocks = (int)std::min<int64_t>(EW_BLOCKS * 4, ceil_div64(std::max<int64_t>(n4, n - n4 * 4), EW_THREADS));
  DSW_CUDA_TRY(launch_pdl(rezero_fwd_kernel, dim3(std::max(blocks, 1)), dim3(EW_THREADS), 0, static_cast<cudaStream_t>(stream), pdl_enabled(),
                          conv_out, skip, w, y, n4, n));
  return check_launch();
}

int dsw_rezero_bwd(const float* g, const float* conv_out, const float* w, float* d_conv_out, float* d_w, void* workspace,
                   size_t workspace_bytes, int64_t n, void* stream) {
  if (!g || !w || n <= 0 || (!d_conv_out && !d_w) || (d_w && !conv_out)) return DSW_ERR_BAD_ARGUMENT;
  if (d_w && (!workspace || workspace_bytes < dsw_rezero_bwd_workspace_bytes())) return DSW_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool v4 = al16(g) && (!conv_out || al16(conv_out)) && (!d_conv_out || al16(d_conv_out));
  const int64_t n4 = v4 ? n / 4 : 0;
  float* partial = d_w ? static_cast<float*>(workspace) : nullptr;
  DSW_CUDA_TRY(launch_pdl(rezero_bwd_kernel, dim3(EW_BLOCKS), dim3(EW_THREADS), 0, st, pdl_enabled(), g, conv_out, w, d_conv_out, partial, n4, n));
  DSW_TRY(check_launch());
  if (d_w) {
    DSW_CUDA_TRY(launch_pdl(rezero_reduce_kernel, dim3(1), dim3(EW_THREADS), 0, st, pdl_enabled(), (const float*)partial, (int32_t)EW_BLOCKS, d_w));
    DSW_TRY(check_launch());
  }
  return DSW_OK;
}

}  // extern "C"
