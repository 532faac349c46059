// Sparse "hop" kernels:  O = alpha * (A . X) + beta * Z + G   on channel-last [B][rows][F] slabs.
//
// One hop of the Chebyshev recurrence x_k = 2 L x_{k-1} - x_{k-2} (reference
// modules/layers.py:163-169), one step of its adjoint, and the pool / unpool remap
// (layers.py:956-964) are all this operation with different (alpha, beta, Z, G).
//
// Two layouts:
//  * hop_rb_kernel  — row-block union panels (dsw_rb): a 16-lane group owns R consecutive rows and a
//    64-feature slab; every gathered float4 feeds R FMAs.  A CTA covers a compact run of rows of
//    one sample so the gathered rows (tile + halo) are served by L1 after the first touch.
//  * hop_csr_kernel — plain CSR, float4 or scalar lanes; used for F % 4 != 0, unaligned strides,
//    and operators whose rows do not overlap (pool matrices).
#include "dsw_internal.cuh"

namespace dsw {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int R>
__global__ void __launch_bounds__(512) hop_rb_kernel(const int32_t* __restrict__ blkptr,
                                                     const int32_t* __restrict__ ucol,
                                                     const float* __restrict__ uval, int32_t n_blocks,
                                                     int32_t n_rows, HopArgs a) {
  // thread -> (row-block slot, float4 lane inside a 64-feature slab)
  const int lane16 = threadIdx.x & 15;
  const int slot = threadIdx.x >> 4;
  const int rb = blockIdx.x * (blockDim.x >> 4) + slot;
  const int c4 = blockIdx.y * 16 + lane16;
  const int b = blockIdx.z;
  if (rb >= n_blocks || c4 * 4 >= a.F) return;

  const float* __restrict__ xb = a.X + b * a.x_sB + c4 * 4;
  float4 acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);

  int u = __ldg(blkptr + rb);
  const int u1 = __ldg(blkptr + rb + 1);
  // 2-way unrolled so two independent gathers are in flight per thread
  for (; u + 2 <= u1; u += 2) {
    const int ca = __ldg(ucol + u), cb = __ldg(ucol + u + 1);
    float wa[R], wb[R];
    if constexpr (R == 4) {
      const float4 t0 = ldg4(uval + (int64_t)u * 4), t1 = ldg4(uval + (int64_t)u * 4 + 4);
      wa[0] = t0.x, wa[1] = t0.y, wa[2] = t0.z, wa[3] = t0.w;
      wb[0] = t1.x, wb[1] = t1.y, wb[2] = t1.z, wb[3] = t1.w;
    } else {
      // u may be odd: only 8-byte alignment is guaranteed for 2-row panels
      const float2 t0 = __ldg(reinterpret_cast<const float2*>(uval + (int64_t)u * 2));
      const float2 t1 = __ldg(reinterpret_cast<const float2*>(uval + (int64_t)u * 2 + 2));
      wa[0] = t0.x, wa[1] = t0.y, wb[0] = t1.x, wb[1] = t1.y;
    }
    const float4 xa = ldg4(xb + ca * a.x_sV);
    const float4 xv = ldg4(xb + cb * a.x_sV);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      acc[r].x = fmaf(wa[r], xa.x, acc[r].x);
      acc[r].y = fmaf(wa[r], xa.y, acc[r].y);
      acc[r].z = fmaf(wa[r], xa.z, acc[r].z);
      acc[r].w = fmaf(wa[r], xa.w, acc[r].w);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      acc[r].x = fmaf(wb[r], xv.x, acc[r].x);
      acc[r].y = fmaf(wb[r], xv.y, acc[r].y);
      acc[r].z = fmaf(wb[r], xv.z, acc[r].z);
      acc[r].w = fmaf(wb[r], xv.w, acc[r].w);
    }
  }
  if (u < u1) {
    const int ca = __ldg(ucol + u);
    const float4 xa = ldg4(xb + ca * a.x_sV);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float w = __ldg(uval + (int64_t)u * R + r);
      acc[r].x = fmaf(w, xa.x, acc[r].x);
      acc[r].y = fmaf(w, xa.y, acc[r].y);
      acc[r].z = fmaf(w, xa.z, acc[r].z);
      acc[r].w = fmaf(w, xa.w, acc[r].w);
    }
  }

#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int row = rb * R + r;
    if (row >= n_rows) break;
    float4 o = make_float4(a.alpha * acc[r].x, a.alpha * acc[r].y, a.alpha * acc[r].z, a.alpha * acc[r].w);
    if (a.Z) {
      const float4 z = ldg4(a.Z + b * a.z_sB + row * a.z_sV + c4 * 4);
      o.x = fmaf(a.beta, z.x, o.x), o.y = fmaf(a.beta, z.y, o.y);
      o.z = fmaf(a.beta, z.z, o.z), o.w = fmaf(a.beta, z.w, o.w);
    }
    if (a.G) {
      const float4 g = ldg4(a.G + b * a.g_sB + row * a.g_sV + c4 * 4);
      o.x += g.x, o.y += g.y, o.z += g.z, o.w += g.w;
    }
    *reinterpret_cast<float4*>(a.O + b * a.o_sB + row * a.o_sV + c4 * 4) = o;
  }
}

// Plain CSR.  LPR (power of two) lanes cooperate on one row; each lane walks columns
// c = lane, lane + LPR, ... in units of VEC floats.
template <int VEC>
__global__ void __launch_bounds__(256) hop_csr_kernel(const int32_t* __restrict__ rowptr,
                                                      const int32_t* __restrict__ col,
                                                      const float* __restrict__ val, int32_t n_rows,
                                                      int32_t lpr_log2, HopArgs a) {
  const int lpr = 1 << lpr_log2;
  const int lane = threadIdx.x & (lpr - 1);
  const int row = blockIdx.x * (blockDim.x >> lpr_log2) + (threadIdx.x >> lpr_log2);
  const int b = blockIdx.y;
  if (row >= n_rows) return;
  const int e0 = __ldg(rowptr + row), e1 = __ldg(rowptr + row + 1);
  const int nvec = a.F / VEC;
  const float* __restrict__ xb = a.X + b * a.x_sB;
  for (int cv = lane; cv < nvec; cv += lpr) {
    const int f = cv * VEC;
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    int e = e0;
    for (; e + 2 <= e1; e += 2) {
      const int ca = __ldg(col + e), cb = __ldg(col + e + 1);
      const float wa = __ldg(val + e), wb = __ldg(val + e + 1);
      if constexpr (VEC == 4) {
        const float4 xa = ldg4(xb + ca * a.x_sV + f), xv = ldg4(xb + cb * a.x_sV + f);
        acc[0] = fmaf(wa, xa.x, acc[0]), acc[1] = fmaf(wa, xa.y, acc[1]);
        acc[2] = fmaf(wa, xa.z, acc[2]), acc[3] = fmaf(wa, xa.w, acc[3]);
        acc[0] = fmaf(wb, xv.x, acc[0]), acc[1] = fmaf(wb, xv.y, acc[1]);
        acc[2] = fmaf(wb, xv.z, acc[2]), acc[3] = fmaf(wb, xv.w, acc[3]);
      } else {
        const float xa = __ldg(xb + ca * a.x_sV + f), xv = __ldg(xb + cb * a.x_sV + f);
        acc[0] = fmaf(wa, xa, acc[0]);
        acc[0] = fmaf(wb, xv, acc[0]);
      }
    }
    if (e < e1) {
      const int ca = __ldg(col + e);
      const float wa = __ldg(val + e);
      if constexpr (VEC == 4) {
        const float4 xa = ldg4(xb + ca * a.x_sV + f);
        acc[0] = fmaf(wa, xa.x, acc[0]), acc[1] = fmaf(wa, xa.y, acc[1]);
        acc[2] = fmaf(wa, xa.z, acc[2]), acc[3] = fmaf(wa, xa.w, acc[3]);
      } else {
        acc[0] = fmaf(wa, __ldg(xb + ca * a.x_sV + f), acc[0]);
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float o = a.alpha * acc[i];
      if (a.Z) o = fmaf(a.beta, __ldg(a.Z + b * a.z_sB + row * a.z_sV + f + i), o);
      if (a.G) o += __ldg(a.G + b * a.g_sB + row * a.g_sV + f + i);
      acc[i] = o;
    }
    float* op = a.O + b * a.o_sB + row * a.o_sV + f;
    if constexpr (VEC == 4) {
      *reinterpret_cast<float4*>(op) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
      op[0] = acc[0];
    }
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static bool vec4_ok(const HopArgs& a) {
  if (a.F % 4) return false;
  if (!aligned16(a.X) || !aligned16(a.O) || (a.x_sB | a.x_sV | a.o_sB | a.o_sV) % 4) return false;
  if (a.Z && (!aligned16(a.Z) || (a.z_sB | a.z_sV) % 4)) return false;
  if (a.G && (!aligned16(a.G) || (a.g_sB | a.g_sV) % 4)) return false;
  return true;
}

int launch_hop(const dsw_csr& A, const dsw_rb& rb, const HopArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.F <= 0 || !a.X || !a.O) return DSW_ERR_BAD_ARGUMENT;
  if (a.B > 65535) return DSW_ERR_UNSUPPORTED;
  const bool v4 = vec4_ok(a);
  if (v4 && rb.R > 0) {
    const int threads = 512;
    const int slots = threads / 16;
    dim3 grid(ceil_div(rb.n_blocks, slots), ceil_div(a.F, 64), a.B);
    if (rb.R == 4)
      hop_rb_kernel<4><<<grid, threads, 0, st>>>(rb.blkptr, rb.ucol, rb.uval, rb.n_blocks, A.n_rows, a);
    else
      hop_rb_kernel<2><<<grid, threads, 0, st>>>(rb.blkptr, rb.ucol, rb.uval, rb.n_blocks, A.n_rows, a);
    return check_launch();
  }
  const int vec = v4 ? 4 : 1;
  const int nvec = a.F / vec;
  int lpr_log2 = 0;
  while ((1 << lpr_log2) < nvec && lpr_log2 < 8) ++lpr_log2;
  const int rows_per_cta = 256 >> lpr_log2;
  dim3 grid(ceil_div(A.n_rows, rows_per_cta), a.B);
  if (v4)
    hop_csr_kernel<4><<<grid, 256, 0, st>>>(A.rowptr, A.col, A.val, A.n_rows, lpr_log2, a);
  else
    hop_csr_kernel<1><<<grid, 256, 0, st>>>(A.rowptr, A.col, A.val, A.n_rows, lpr_log2, a);
  return check_launch();
}

}  // namespace dsw

using namespace dsw;

extern "C" {

int dsw_spmm_fwd(const dsw_plan* mat, const float* x, int64_t x_sB, int64_t x_sV, float* y, int32_t B,
                 int32_t F, void* stream) {
  if (!mat || !x || !y || B <= 0 || F <= 0) return DSW_ERR_BAD_ARGUMENT;
  HopArgs a;
  a.X = x, a.x_sB = x_sB, a.x_sV = x_sV;
  a.O = y, a.o_sV = F, a.o_sB = (int64_t)mat->fwd.n_rows * F;
  a.B = B, a.F = F;
  return launch_hop(mat->fwd, mat->fwd_rb, a, static_cast<cudaStream_t>(stream));
}

int dsw_spmm_bwd(const dsw_plan* mat, const float* dy, int64_t dy_sB, int64_t dy_sV, float* dx, int32_t B,
                 int32_t F, void* stream) {
  if (!mat || !dy || !dx || B <= 0 || F <= 0) return DSW_ERR_BAD_ARGUMENT;
  HopArgs a;
  a.X = dy, a.x_sB = dy_sB, a.x_sV = dy_sV;
  a.O = dx, a.o_sV = F, a.o_sB = (int64_t)mat->tr.n_rows * F;
  a.B = B, a.F = F;
  return launch_hop(mat->tr, mat->tr_rb, a, static_cast<cudaStream_t>(stream));
}

int dsw_cheb_terms(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, float* terms, int32_t B,
                   int32_t F, int32_t K, void* stream) {
  if (!lap || !x || B <= 0 || F <= 0 || K < 1) return DSW_ERR_BAD_ARGUMENT;
  if (K > DSW_MAX_K) return DSW_ERR_UNSUPPORTED;
  if (lap->fwd.n_rows != lap->fwd.n_cols) return DSW_ERR_SHAPE;
  if (K > 1 && !terms) return DSW_ERR_BAD_ARGUMENT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t V = lap->fwd.n_rows, plane = (int64_t)B * V * F;
  for (int k = 1; k < K; ++k) {
    HopArgs a;
    a.B = B, a.F = F;
    a.O = terms + (k - 1) * plane, a.o_sB = V * F, a.o_sV = F;
    if (k == 1) {
      a.X = x, a.x_sB = x_sB, a.x_sV = x_sV;
    } else {
      a.X = terms + (k - 2) * plane, a.x_sB = V * F, a.x_sV = F;
      a.alpha = 2.f, a.beta = -1.f;
      if (k == 2) {
        a.Z = x, a.z_sB = x_sB, a.z_sV = x_sV;
      } else {
        a.Z = terms + (k - 3) * plane, a.z_sB = V * F, a.z_sV = F;
      }
    }
    DSW_TRY(launch_hop(lap->fwd, lap->fwd_rb, a, st));
  }
  return DSW_OK;
}

}  // extern "C"
