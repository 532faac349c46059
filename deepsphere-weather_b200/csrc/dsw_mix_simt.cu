// CUDA-core fp32 dense kernels: the (K*Fin)->Fout channel mix of conv_cheb (reference
// modules/layers.py:171-177), its transpose (dy . W^T) for the input gradient, and the weight
// gradient reduction.  Exact fp32 FMA arithmetic; this is the always-available path and the one
// used for shapes the tcgen05 kernel does not take.
#include <algorithm>

#include "dsw_internal.cuh"

namespace dsw {

namespace {
constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4, APAD = 4;
}

__device__ __forceinline__ int64_t row_offset(int64_t n, int32_t V, int64_t sB, int64_t sV) {
  const int64_t b = n / V;
  return b * sB + (n - b * V) * sV;
}

template <bool VEC4>
__global__ void __launch_bounds__(256) mix_simt_kernel(MixArgs a) {
  __shared__ __align__(16) float As[BK][BM + APAD];
  __shared__ __align__(16) float Bs[BK][BN];
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const int64_t n0 = (int64_t)blockIdx.x * BM;
  const int c0 = blockIdx.y * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // per-thread A-load coordinates (2 x float4 along the reduction dim)
  int a_row[2], a_kq[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = t + i * 256;
    a_row[i] = idx >> 2;
    a_kq[i] = (idx & 3) * 4;
  }

  for (int p = 0; p < a.P; ++p) {
    const float* __restrict__ Ap = a.A[p];
    int64_t roff[2];
    bool rok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t n = n0 + a_row[i];
      rok[i] = n < a.N;
      roff[i] = rok[i] ? row_offset(n, a.rows_per_batch, a.a_sB[p], a.a_sV[p]) : 0;
    }
    for (int k0 = 0; k0 < a.Ka; k0 += BK) {
      // ---- A tile -> As[k][m] (transposed) ----
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int kk = k0 + a_kq[i];
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (rok[i]) {
          if (VEC4 && kk + 3 < a.Ka) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(Ap + roff[i] + kk));
            v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (kk + j < a.Ka) v[j] = __ldg(Ap + roff[i] + kk + j);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) As[a_kq[i] + j][a_row[i]] = v[j];
      }
      // ---- B tile -> Bs[k][c] ----
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = t + i * 256;
        const int kk = idx >> 6, c = idx & 63;
        float v = 0.f;
        const int cg = c0 + c;
        if (k0 + kk < a.Ka && cg < a.Nc) {
          const int cp = cg / a.Cw, cc = cg - cp * a.Cw;
          v = __ldg(a.Bm + p * a.sBp + (int64_t)(k0 + kk) * a.sBk + cp * a.sBc1 + cc * a.sBc0);
        }
        Bs[kk][c] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * TN]);
        const float av[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[TN] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue: bias, activation, store ----
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int cg = c0 + tx * TN + j;
    if (cg >= a.Nc) continue;
    const float bz = (a.bias && cg < a.bias_n) ? __ldg(a.bias + cg) : 0.f;
    const int cp = cg / a.Cw, cc = cg - cp * a.Cw;
    float* __restrict__ Cb = a.C + cp * a.sCp + cc;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int64_t n = n0 + ty * TM + i;
      if (n >= a.N) continue;
      float v = acc[i][j] + bz;
      if (a.R) {
        const float rr = __ldg(a.R + n * a.ldr + cg);
        v = a.r_mode == 1 ? (rr > 0.f ? v : 0.f) : fmaf(a.r_scale ? __ldg(a.r_scale) : 1.f, rr, v);
      }
      if (a.act == 1) v = fmaxf(v, 0.f);
      Cb[n * a.ldc] = v;
    }
  }
}

int launch_mix_simt(const MixArgs& a, cudaStream_t st) {
  bool vec4 = (a.Ka % 4 == 0);
  for (int p = 0; p < a.P && vec4; ++p)
    vec4 = ((reinterpret_cast<uintptr_t>(a.A[p]) & 15) == 0) && (a.a_sB[p] % 4 == 0) && (a.a_sV[p] % 4 == 0);
  dim3 grid((unsigned)ceil_div64(a.N, BM), ceil_div(a.Nc, BN));
  if (vec4)
    mix_simt_kernel<true><<<grid, 256, 0, st>>>(a);
  else
    mix_simt_kernel<false><<<grid, 256, 0, st>>>(a);
  return check_launch();
}

// ---------------------------------------------------------------------------------------------------
// Weight gradient.  grid = (K * ceil(Fin/64) + 1, ceil(Fout/64), nsplit); the extra x-tile feeds a
// row of ones so that dbias falls out of the same reduction.  Each CTA reduces its row range into a
// 64x64 register tile and writes a partial; wgrad_reduce_kernel sums partials in a fixed order.
// ---------------------------------------------------------------------------------------------------
namespace {
constexpr int WT = 64, WR = 32;
}

__global__ void __launch_bounds__(256) wgrad_simt_kernel(WgradArgs a, int32_t ftiles, int64_t rows_per_split) {
  __shared__ __align__(16) float As[WR][WT];
  __shared__ __align__(16) float Bs[WR][WT];
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int mt = blockIdx.x;
  const bool bias_tile = (mt == a.Ka * ftiles);
  const int k = bias_tile ? 0 : mt / ftiles;
  const int f0 = bias_tile ? 0 : (mt - k * ftiles) * WT;
  const int otiles = (a.Fout + WT - 1) / WT;
  const int kb_plane = blockIdx.y / otiles;  // Fout-side plane
  const int o0 = (blockIdx.y - kb_plane * otiles) * WT;
  const float* __restrict__ Yp = a.Y[kb_plane];
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_end = (r_begin + rows_per_split < a.N) ? r_begin + rows_per_split : a.N;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const float* __restrict__ Tk = a.T[k];
  const int64_t sB = a.t_sB[k], sV = a.t_sV[k];

  for (int64_t r0 = r_begin; r0 < r_end; r0 += WR) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = t + i * 256;
      const int rr = idx >> 6, cc = idx & 63;
      const int64_t n = r0 + rr;
      float va = 0.f, vb = 0.f;
      if (n < r_end) {
        if (bias_tile)
          va = (cc == 0) ? 1.f : 0.f;
        else if (f0 + cc < a.Fin)
          va = __ldg(Tk + row_offset(n, a.rows_per_batch, sB, sV) + f0 + cc);
        if (o0 + cc < a.Fout) vb = __ldg(Yp + n * a.Fout + o0 + cc);
      }
      As[rr][cc] = va;
      Bs[rr][cc] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < WR; ++rr) {
      const float4 av = *reinterpret_cast<const float4*>(&As[rr][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[rr][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int64_t prow = (int64_t)a.K * a.Fin + 1;
  float* __restrict__ P = a.partial + (int64_t)blockIdx.z * prow * a.Fout;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int fl = ty * 4 + i;
    int64_t m;
    if (bias_tile) {
      if (fl != 0 || kb_plane != 0) continue;
      m = prow - 1;
    } else {
      if (f0 + fl >= a.Fin) continue;
      m = (int64_t)(f0 + fl) * a.K + k + kb_plane;  // reference weight layout [Fin][K][Fout]
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = o0 + tx * 4 + j;
      if (o < a.Fout) P[m * a.Fout + o] = acc[i][j];
    }
  }
}

// Sum of the split partials in a fixed order (deterministic).  A CTA owns 32 consecutive outputs; its 8
// warps each add every 8th partial (coalesced 128-byte reads, independent loads in flight), then warp 0
// adds the 8 warp sums in warp order.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int32_t nsplit, int64_t n_w,
                                                           int32_t Fout, float* __restrict__ dW, float* __restrict__ dbias) {
  __shared__ float red[8][32];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t total = n_w + Fout;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  float s = 0.f;
  if (i < total) {
    const float* p = partial + i;
    int sp = w;
    for (; sp + 24 < nsplit; sp += 32) {
      const float v0 = __ldg(p + (int64_t)sp * total), v1 = __ldg(p + (int64_t)(sp + 8) * total);
      const float v2 = __ldg(p + (int64_t)(sp + 16) * total), v3 = __ldg(p + (int64_t)(sp + 24) * total);
      s += v0, s += v1, s += v2, s += v3;
    }
    for (; sp < nsplit; sp += 8) s += __ldg(p + (int64_t)sp * total);
  }
  red[w][lane] = s;
  __syncthreads();
  if (w == 0 && i < total) {
    float t = red[0][lane];
#pragma unroll
    for (int j = 1; j < 8; ++j) t += red[j][lane];
    if (i < n_w)
      dW[i] = t;
    else if (dbias)
      dbias[i - n_w] = t;
  }
}

int wgrad_pick_nsplit(int64_t N, int32_t Ka, int32_t Kb, int32_t Fin, int32_t Fout) {
  const int64_t tiles = ((int64_t)Ka * ceil_div(Fin, WT) + 1) * Kb * ceil_div(Fout, WT);
  int64_t ns = ceil_div64(148 * 4, tiles);           // ~4 CTAs per SM
  ns = std::min<int64_t>(ns, ceil_div64(N, 4 * WR)); // at least 4 row steps per split
  return (int)std::max<int64_t>(1, std::min<int64_t>(ns, 1024));
}

int launch_wgrad_reduce(const float* partial, int32_t nsplit, int32_t K, int32_t Fin, int32_t Fout, float* dW,
                        float* dbias, cudaStream_t st) {
  const int64_t n_w = (int64_t)K * Fin * Fout;
  const int64_t total = n_w + Fout;
  DSW_CUDA_TRY(launch_pdl(wgrad_reduce_kernel, dim3((unsigned)ceil_div64(total, 32)), dim3(256), 0, st, pdl_enabled(), partial, nsplit, n_w, Fout, dW,
                          dbias));
  return check_launch();
}

// Writes a.nsplit partials at a.partial; the caller sums them with launch_wgrad_reduce.
int launch_wgrad_simt(const WgradArgs& a, cudaStream_t st) {
  const int ftiles = ceil_div(a.Fin, WT);
  int64_t rps = ceil_div64(a.N, a.nsplit);
  rps = ceil_div64(rps, WR) * WR;
  dim3 grid(a.Ka * ftiles + 1, a.Kb * ceil_div(a.Fout, WT), a.nsplit);
  wgrad_simt_kernel<<<grid, 256, 0, st>>>(a, ftiles, rps);
  return check_launch();
}

}  // namespace dsw
