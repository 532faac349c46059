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
#include "dsw_tmap.cuh"

namespace dsw {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// L2-only (cache-global) load: streaming operands that are read once per kernel
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

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
    if (a.act) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
    if (a.M) {  // ReLU mask of a gradient by the ReLU's output (indexed like the output rows)
      const float4 m = ldg4(a.M + b * a.m_sB + row * a.m_sV + c4 * 4);
      o.x = m.x > 0.f ? o.x : 0.f, o.y = m.y > 0.f ? o.y : 0.f, o.z = m.z > 0.f ? o.z : 0.f, o.w = m.w > 0.f ? o.w : 0.f;
    }
    *reinterpret_cast<float4*>(a.O + b * a.o_sB + row * a.o_sV + c4 * 4) = o;
  }
}

// Four-channel planes (the 2-channel output head padded to 4, its gradient): one lane owns a row for NB samples, so the
// row's column / weight list — read uncoalesced, lane by lane — is fetched once per NB samples instead of once per sample
// (the plain CSR kernel below was bound by exactly those sector loads: 34 us per hop on 12 288 nodes x 32 samples).
// Same operation order per output element as hop_csr_kernel (results are bit-identical).
template <int NB>
__global__ void __launch_bounds__(256) hop_csr_f4_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                         const float* __restrict__ val, int32_t n_rows, HopArgs a) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 256 + threadIdx.x;
  const int b0 = blockIdx.y * NB;
  if (row >= n_rows) return;
  const int e0 = __ldg(rowptr + row), e1 = __ldg(rowptr + row + 1);
  const int nb = min(NB, a.B - b0);
  const float* __restrict__ xb = a.X + (int64_t)b0 * a.x_sB;
  float4 acc[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = e0; e < e1; ++e) {
    const int c = __ldg(col + e);
    const float w = __ldg(val + e);
    float4 x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) x[i] = i < nb ? ldg4(xb + i * a.x_sB + c * a.x_sV) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      acc[i].x = fmaf(w, x[i].x, acc[i].x), acc[i].y = fmaf(w, x[i].y, acc[i].y);
      acc[i].z = fmaf(w, x[i].z, acc[i].z), acc[i].w = fmaf(w, x[i].w, acc[i].w);
    }
  }
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    if (i >= nb) break;
    const int64_t b = b0 + i;
    float o[4] = {a.alpha * acc[i].x, a.alpha * acc[i].y, a.alpha * acc[i].z, a.alpha * acc[i].w};
    if (a.Z) {
      const float4 z = ldg4(a.Z + b * a.z_sB + row * a.z_sV);
      o[0] = fmaf(a.beta, z.x, o[0]), o[1] = fmaf(a.beta, z.y, o[1]), o[2] = fmaf(a.beta, z.z, o[2]), o[3] = fmaf(a.beta, z.w, o[3]);
    }
    if (a.G) {
      const float4 g = ldg4(a.G + b * a.g_sB + row * a.g_sV);
      o[0] += g.x, o[1] += g.y, o[2] += g.z, o[3] += g.w;
    }
    if (a.act) o[0] = fmaxf(o[0], 0.f), o[1] = fmaxf(o[1], 0.f), o[2] = fmaxf(o[2], 0.f), o[3] = fmaxf(o[3], 0.f);
    if (a.M) {
      const float4 m = ldg4(a.M + b * a.m_sB + row * a.m_sV);
      o[0] = m.x > 0.f ? o[0] : 0.f, o[1] = m.y > 0.f ? o[1] : 0.f, o[2] = m.z > 0.f ? o[2] : 0.f, o[3] = m.w > 0.f ? o[3] : 0.f;
    }
    *reinterpret_cast<float4*>(a.O + b * a.o_sB + row * a.o_sV) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// Plain CSR.  LPR (power of two) lanes cooperate on one row; each lane walks columns
// c = lane, lane + LPR, ... in units of VEC floats.
template <int VEC>
__global__ void __launch_bounds__(256) hop_csr_kernel(const int32_t* __restrict__ rowptr,
                                                      const int32_t* __restrict__ col,
                                                      const float* __restrict__ val, int32_t n_rows,
                                                      int32_t lpr_log2, HopArgs a) {
  pdl_trigger();
  pdl_wait();
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
    // eight entries per trip: all column / weight loads, then all gathers, are in flight together (the 2-entry form below
    // kept two gathers in flight per lane: latency-bound on narrow planes); the FMAs keep the entry order
    if constexpr (VEC == 4) {
      for (; e + 8 <= e1; e += 8) {
        int c[8];
        float w[8];
        float4 xv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = __ldg(col + e + i), w[i] = __ldg(val + e + i);
#pragma unroll
        for (int i = 0; i < 8; ++i) xv[i] = ldg4(xb + c[i] * a.x_sV + f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[0] = fmaf(w[i], xv[i].x, acc[0]), acc[1] = fmaf(w[i], xv[i].y, acc[1]);
          acc[2] = fmaf(w[i], xv[i].z, acc[2]), acc[3] = fmaf(w[i], xv[i].w, acc[3]);
        }
      }
    }
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
      if (a.act) o = fmaxf(o, 0.f);
      if (a.M) o = __ldg(a.M + b * a.m_sB + row * a.m_sV + f + i) > 0.f ? o : 0.f;
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

// ---------------------------------------------------------------------------------------------
// hop_tile_kernel — the production hop for row-block (R = 4) operators.
//
// CTA = TB consecutive row-blocks (128 rows) x one 64-channel slab x `ns` samples.  The tile's slice
// of the union panels (column list -> pre-multiplied byte offsets, R x U weights) is staged in shared
// memory once and reused for every sample, so the inner loop issues two broadcast LDS (offset,
// weights) per union entry and NF4 gathers of 16 bytes that each feed R packed FMAs (FFMA2).
// 8 lanes own one row-block: lane l reads float4 columns l and l + 8 of the slab, i.e. the 8 lanes
// of a row-block touch one full 128-byte line per gather instruction.
// ---------------------------------------------------------------------------------------------
constexpr int TILE_BLOCKS = 32;   // row-blocks per CTA
constexpr int TILE_THREADS = 256; // 8 lanes per row-block

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& x) {
  const float2 ww = make_float2(w, w);
  float2 lo = __ffma2_rn(ww, make_float2(x.x, x.y), make_float2(acc.x, acc.y));
  float2 hi = __ffma2_rn(ww, make_float2(x.z, x.w), make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}

template <int NF4>
__global__ void __launch_bounds__(TILE_THREADS, 3) hop_tile_kernel(const int32_t* __restrict__ blkptr,
                                                                    const int32_t* __restrict__ ucol,
                                                                    const float4* __restrict__ uval, int32_t n_blocks,
                                                                    int32_t n_rows, int32_t ns, HopArgs a) {
  extern __shared__ __align__(256) uint8_t tile_smem[];
  const int tid = threadIdx.x;
  const int blk0 = blockIdx.x * TILE_BLOCKS;
  const int blk1 = min(blk0 + TILE_BLOCKS, n_blocks);
  const int e0 = __ldg(blkptr + blk0);
  const int ne = __ldg(blkptr + blk1) - e0;
  float4* s_val = reinterpret_cast<float4*>(tile_smem);
  uint32_t* s_off = reinterpret_cast<uint32_t*>(tile_smem + (size_t)((ne + 3) & ~3) * 16);
  const uint32_t row_bytes = (uint32_t)a.x_sV * 4u;
  for (int i = tid; i < ne; i += TILE_THREADS) {
    s_val[i] = __ldg(uval + e0 + i);
    s_off[i] = (uint32_t)__ldg(ucol + e0 + i) * row_bytes;
  }
  __syncthreads();

  const int slot = tid >> 3, l8 = tid & 7;
  const int blk = blk0 + slot;
  if (blk >= n_blocks) return;
  const int u0 = __ldg(blkptr + blk) - e0, u1 = __ldg(blkptr + blk + 1) - e0;
  const int c4 = blockIdx.y * (8 * NF4) + l8;  // first float4 column of this lane
  const int nf4 = a.F >> 2;
  bool ok[NF4];
#pragma unroll
  for (int j = 0; j < NF4; ++j) ok[j] = (c4 + 8 * j) < nf4;
  if (!ok[0]) return;

  const int b_begin = blockIdx.z * ns, b_end = min(b_begin + ns, a.B);
  for (int b = b_begin; b < b_end; ++b) {
    const char* __restrict__ xb = reinterpret_cast<const char*>(a.X + (int64_t)b * a.x_sB + c4 * 4);
    float4 acc[4][NF4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < NF4; ++j) acc[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);

    int u = u0;
#pragma unroll 1
    for (; u + 2 <= u1; u += 2) {
      const uint32_t oa = s_off[u], ob = s_off[u + 1];
      const float4 wa = s_val[u], wb = s_val[u + 1];
      float4 xa[NF4], xv[NF4];
#pragma unroll
      for (int j = 0; j < NF4; ++j) {
        xa[j] = ok[j] ? ldg4(reinterpret_cast<const float*>(xb + oa + 128 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
        xv[j] = ok[j] ? ldg4(reinterpret_cast<const float*>(xb + ob + 128 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < NF4; ++j) {
        fma4(acc[0][j], wa.x, xa[j]), fma4(acc[1][j], wa.y, xa[j]);
        fma4(acc[2][j], wa.z, xa[j]), fma4(acc[3][j], wa.w, xa[j]);
      }
#pragma unroll
      for (int j = 0; j < NF4; ++j) {
        fma4(acc[0][j], wb.x, xv[j]), fma4(acc[1][j], wb.y, xv[j]);
        fma4(acc[2][j], wb.z, xv[j]), fma4(acc[3][j], wb.w, xv[j]);
      }
    }
    if (u < u1) {
      const uint32_t oa = s_off[u];
      const float4 wa = s_val[u];
#pragma unroll
      for (int j = 0; j < NF4; ++j) {
        if (!ok[j]) continue;
        const float4 xa = ldg4(reinterpret_cast<const float*>(xb + oa + 128 * j));
        fma4(acc[0][j], wa.x, xa), fma4(acc[1][j], wa.y, xa);
        fma4(acc[2][j], wa.z, xa), fma4(acc[3][j], wa.w, xa);
      }
    }

#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = blk * 4 + r;
      if (row >= n_rows) break;
#pragma unroll
      for (int j = 0; j < NF4; ++j) {
        if (!ok[j]) continue;
        const int64_t col = (int64_t)(c4 + 8 * j) * 4;
        float4 o = make_float4(a.alpha * acc[r][j].x, a.alpha * acc[r][j].y, a.alpha * acc[r][j].z,
                               a.alpha * acc[r][j].w);
        if (a.Z) {
          const float4 z = ldg4(a.Z + b * a.z_sB + row * a.z_sV + col);
          o.x = fmaf(a.beta, z.x, o.x), o.y = fmaf(a.beta, z.y, o.y);
          o.z = fmaf(a.beta, z.z, o.z), o.w = fmaf(a.beta, z.w, o.w);
        }
        if (a.G) {
          const float4 g = ldg4(a.G + b * a.g_sB + row * a.g_sV + col);
          o.x += g.x, o.y += g.y, o.z += g.z, o.w += g.w;
        }
        if (a.act) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
        if (a.M) {  // ReLU mask of a gradient by the ReLU's output (indexed like the output rows)
          const float4 m = ldg4(a.M + b * a.m_sB + row * a.m_sV + col);
          o.x = m.x > 0.f ? o.x : 0.f, o.y = m.y > 0.f ? o.y : 0.f, o.z = m.z > 0.f ? o.z : 0.f, o.w = m.w > 0.f ? o.w : 0.f;
        }
        *reinterpret_cast<float4*>(a.O + b * a.o_sB + row * a.o_sV + col) = o;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// hop_team_kernel — the production hop for row-block (R = 4) operators.
//
// CTA = one tile of DSW_TILE_BLOCKS consecutive row-blocks (128 rows).  The tile's entry-major
// padded union panels (R weights + the local source-row offset per entry step and row-block) and its
// source-row list are staged in shared memory once per CTA.  Up to three independent 128-thread
// *teams* then walk the CTA's work items (sample, 64-channel slab): a team stages the distinct source
// rows the tile gathers for its item (own rows + halo; 16-byte cp.async.cg straight from L2, ~64 KB
// in flight per team, no registers), computes the neighbour sums entirely out of shared memory and
// writes the 128 output rows.  Teams synchronise only among themselves (named barriers), so one
// team's staging overlaps the other teams' arithmetic and the memory latency is decoupled from the
// FMA loop.
//
// Inside a team 4 lanes own one row-block; a lane holds the 4 rows x 4 float4 columns accumulator
// tile, so every 16-byte shared-memory read feeds R = 4 packed FMAs (FFMA2) and the per-step
// broadcast reads (offset, weights) are amortised over 64 FMAs per lane.  The entry loop is software
// pipelined by hand (offsets two steps ahead, weights and source values one step ahead).  The two
// row-blocks that share a quarter-warp phase read opposite 64-byte halves of their rows (column
// group XOR parity), which keeps the 128-bit reads bank-conflict free.
// ---------------------------------------------------------------------------------------------
constexpr int MAX_TEAMS = 5;
constexpr int PANEL_PAD = DSW_PANEL_PAD;  // zero entry steps appended so that the pipeline may over-read

struct TeamHopPlan {
  const int32_t* blkptr;
  const int32_t* tp_ptr;
  const float4* tp_val;
  const uint32_t* tp_off;
  const int32_t* tile_ptr;
  const int32_t* tile_row;
  const int32_t* tpc_ptr;
  const int32_t* tpc_row;
  const uint32_t* tpc_meta;
  int32_t cap_pieces;
  int32_t n_blocks, n_rows, cap_len, cap_rows;
  int32_t n_slabs;        // ceil(F / 64)
  int32_t n_items;        // B * n_slabs
  int32_t items_per_cta;
  int32_t n_teams;
  int32_t* cnt;           // per-tile claim counters of this launch (dynamic item scheduling) or null
  int32_t n_ctas;         // CTAs of this launch (the last one to leave zeroes the counters again)
  int32_t prefetch_zg;    // 1 = L2-prefetch the next item's Z / G rows when its tile transfer is issued
  int32_t pdl;            // 1 = launched with programmatic stream serialisation: the prologue may overlap the previous kernel
  const int32_t* perm;    // PERM kernels: original row id of every permuted row position ([n_blocks * 4], -1 = padding)
  int32_t debug_skip;     // timing experiments only: 1 = skip staging, 2 = skip the entry loop
};

// Phase timing for tuning (DSW_OPT_DEBUG = 4): cycles summed over teams, read back by dsw_debug_counters.
__device__ unsigned long long g_hop_prof[8];

struct HopMaps {
  CUtensorMap m[8];  // box rows 1, 2, 4, .. 128 over the gather source [B][n_cols][F]
};

__device__ __forceinline__ uint32_t hop_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hop_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void hop_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hop_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

template <int NT>
__device__ __forceinline__ void team_sync(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(NT) : "memory");
}

template <int NJ>
__device__ __forceinline__ void fma_step(float4 (&acc)[4][NJ], const float4& w, const float4 (&x)[NJ]) {
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    fma4(acc[0][j], w.x, x[j]);
    fma4(acc[1][j], w.y, x[j]);
    fma4(acc[2][j], w.z, x[j]);
    fma4(acc[3][j], w.w, x[j]);
  }
}

// L2 prefetch of the Z / G rows of one work item (kept out of line so that it does not disturb the
// register allocation of the entry loop).
__device__ __noinline__ void hop_prefetch_zg(const float* z, int64_t z_sV, const float* g, int64_t g_sV, int rows, bool two_lines,
                                             int tt, int team_threads) {
  for (int r = tt; r < rows; r += team_threads) {
    if (z != nullptr) {
      const float* p = z + r * z_sV;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      if (two_lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 32));
    }
    if (g != nullptr) {
      const float* p = g + r * g_sV;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      if (two_lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 32));
    }
  }
}

// LPR = lanes per row-block: 4 (a lane holds 4 rows x 16 channels) or 8 (4 rows x 8 channels: twice the
// warps per team and half the registers per lane).
// PERM: the plan tiles a locality-preserving permutation of the rows (orderings without locality, e.g.
// row-major lat-lon grids): row-block position p holds original row P.perm[p].
template <bool TMA, int LPR, bool PERM>
__global__ void __launch_bounds__(DSW_TILE_BLOCKS* LPR* MAX_TEAMS, 1)
    hop_team_kernel(const TeamHopPlan P, const HopArgs a, const __grid_constant__ HopMaps maps) {
  extern __shared__ __align__(256) uint8_t tile_smem[];
  constexpr int TEAM_THREADS = DSW_TILE_BLOCKS * LPR;
  constexpr int NJ = 16 / LPR;  // float4 columns per lane
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int blk0 = tile * DSW_TILE_BLOCKS;
  const int t0 = __ldg(P.tp_ptr + tile);
  const int len = __ldg(P.tp_ptr + tile + 1) - t0 - PANEL_PAD;  // the plan's panels carry the pad steps
  const int r0 = __ldg(P.tile_ptr + tile);
  const int nrows = __ldg(P.tile_ptr + tile + 1) - r0;
  // Programmatic dependent launch: the next kernel of the stream may start its CTAs (plan staging only) as
  // soon as every CTA of this grid has got this far, i.e. while our last wave is still computing.
  if (P.pdl) pdl_trigger();

  // shared memory: [staged rows: n_teams x cap_rows x 256 B | weights | offsets | source-row ids]
  const size_t xbuf_bytes = (size_t)P.cap_rows * 256;
  const int cap_steps = P.cap_len + PANEL_PAD;
  float4* s_val = reinterpret_cast<float4*>(tile_smem + (size_t)P.n_teams * xbuf_bytes);
  uint32_t* s_off = reinterpret_cast<uint32_t*>(s_val + (size_t)cap_steps * DSW_TILE_BLOCKS);
  int32_t* s_row = reinterpret_cast<int32_t*>(s_off + (size_t)cap_steps * DSW_TILE_BLOCKS);
  // TMA variant: s_row holds the pieces (first source row), s_meta their (local row, log2 length)
  uint32_t* s_meta = reinterpret_cast<uint32_t*>(s_row + ((P.cap_rows + 1) & ~1));
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_meta + ((P.cap_pieces + 1) & ~1));
  int32_t* s_claim = reinterpret_cast<int32_t*>(s_bar + MAX_TEAMS);  // [team][2] next claimed item; [2 * MAX_TEAMS] = exit flag
  const int pc0 = TMA ? __ldg(P.tpc_ptr + tile) : 0;
  const int npieces = TMA ? __ldg(P.tpc_ptr + tile + 1) - pc0 : 0;

  // Dynamic scheduling: the CTAs of one tile (blockIdx.y = 0 .. gridDim.y - 1) claim its work items
  // from a global counter, so late CTAs share what is left and CTAs that find nothing leave at once.
  int32_t* cnt = P.cnt ? P.cnt + tile : nullptr;
  // Leaving protocol: the counters must be zero again for the next launch that uses this set (stream
  // order, CUDA-graph replays included).  Every CTA counts itself out on cnt[n_tiles]; the last one —
  // nobody claims any more — zeroes the whole set.
  auto leave = [&]() {
    __threadfence();
    const int32_t gone = atomicAdd(P.cnt + gridDim.x, 1);
    if (gone == P.n_ctas - 1) {
      for (int t = 0; t <= (int)gridDim.x; ++t) P.cnt[t] = 0;
    }
  };
  if (cnt) {
    if (tid == 0) s_claim[2 * MAX_TEAMS] = atomicAdd(cnt, 0) >= P.n_items;
    __syncthreads();
    if (s_claim[2 * MAX_TEAMS]) {
      if (tid == 0) leave();
      return;
    }
  }

  // ---- prologue, part 1: the source-row list (TMA: piece list + mbarriers) ----
  if (TMA) {
    for (int i = tid; i < npieces; i += blockDim.x) {
      s_row[i] = __ldg(P.tpc_row + pc0 + i);
      s_meta[i] = __ldg(P.tpc_meta + pc0 + i);
    }
    if (tid < MAX_TEAMS) hop_mbar_init(hop_smem_u32(s_bar + tid), 1);
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else {
    for (int i = tid; i < nrows; i += blockDim.x) s_row[i] = __ldg(P.tile_row + r0 + i);
  }
  __syncthreads();

  const int team = tid / TEAM_THREADS;
  const int tt = tid - team * TEAM_THREADS;
  uint8_t* xs = tile_smem + (size_t)(team < P.n_teams ? team : 0) * xbuf_bytes;
  const uint32_t xs_u32 = (uint32_t)__cvta_generic_to_shared(xs);
  const int item_begin = cnt ? 0 : blockIdx.y * P.items_per_cta;
  const int item_end = cnt ? P.n_items : min(item_begin + P.items_per_cta, P.n_items);
  // Stages the source rows of one item into the team's buffer.  TMA: issued by the lanes of the team's
  // first warp; called for the first item right here (the transfer overlaps the panel staging below) and
  // for the next item as soon as the entry loop of the current one is done (the buffer is free then), so
  // that the transfer overlaps the output stores and the next item's Z / G loads.
  auto stage = [&](int item) {
    const int b = item / P.n_slabs, slab = item - b * P.n_slabs;
    const int slab_f = min(64, a.F - slab * 64);
    const int cpr = slab_f >> 2;
    (void)cpr;
    if (TMA && P.debug_skip == 1) {
    } else if (TMA) {
      // one lane issues a tensor-map box per piece (run of consecutive source rows); the copies bypass
      // registers and L1, complete on the team's mbarrier, and out-of-range channels are zero-filled
      {
        // The boxes are dealt round-robin over the team's warps: a TMA instruction costs ~60 cycles of issue from one
        // warp, and the issuing warps are compute warps — one warp issuing all ~24 boxes ran 1.4 k cycles behind its peer
        // on every item.  (The phase cannot complete before lane 0's expect_tx arrival, so boxes may be issued ahead of it.)
        constexpr int NW = TEAM_THREADS / 32;
        const uint32_t bar = hop_smem_u32(s_bar + team);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (tt == 0) hop_mbar_expect_tx(bar, (uint32_t)nrows * 256u);
        for (int i = (tt & 31) * NW + (tt >> 5); i < npieces; i += 32 * NW) {
          const uint32_t meta = s_meta[i];
          tma_load_3d(xs_u32 + (meta >> 8) * 256u, &maps.m[meta & 7u], slab * 64, s_row[i], b, bar);
        }
      }
      // pull the item's Z / G rows (read at the top of the item, straight into the accumulators) towards
      // L2 now, while the current item's stores and the tile transfer are in flight
      if (!PERM && P.prefetch_zg && (a.Z != nullptr || a.G != nullptr)) {
        const int64_t row0 = (int64_t)blk0 * 4;
        const int rows = (int)min((int64_t)(4 * DSW_TILE_BLOCKS), (int64_t)P.n_rows - row0);
        hop_prefetch_zg(a.Z ? a.Z + b * a.z_sB + row0 * a.z_sV + slab * 64 : nullptr, a.z_sV,
                        a.G ? a.G + b * a.g_sB + row0 * a.g_sV + slab * 64 : nullptr, a.g_sV, rows, slab_f > 32, tt, TEAM_THREADS);
      }
    } else {
      const float* xb = a.X + (int64_t)b * a.x_sB + slab * 64;
      const int c = tt & 15;
      if (c < cpr) {
        const float* xc = xb + c * 4;
        const uint32_t dc = xs_u32 + (uint32_t)c * 16u;
#pragma unroll 4
        for (int r = tt >> 4; r < nrows; r += TEAM_THREADS / 16) {
          const float* src = xc + (int64_t)s_row[r] * a.x_sV;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dc + (uint32_t)r * 256u), "l"(src) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };
  // first item of the team: claimed (dynamic) or item_begin + team (static)
  int item = item_begin + team;
  if (cnt && team < P.n_teams) {
    if (tt == 0) s_claim[2 * team] = atomicAdd(cnt, 1);
    team_sync<TEAM_THREADS>(team);
    item = s_claim[2 * team];
  }
  if (TMA && !P.pdl && team < P.n_teams && item < item_end) stage(item);

  // ---- prologue, part 2: the entry-major weight / offset panels ----
  {
    const float4* gv = P.tp_val + (size_t)t0 * DSW_TILE_BLOCKS;
    const uint32_t* go = P.tp_off + (size_t)t0 * DSW_TILE_BLOCKS;
    const int n_real = len * DSW_TILE_BLOCKS, n_all = (len + PANEL_PAD) * DSW_TILE_BLOCKS;
    for (int i = tid; i < n_all; i += blockDim.x) {
      s_val[i] = i < n_real ? __ldg(gv + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      s_off[i] = i < n_real ? __ldg(go + i) : 0u;
    }
  }
  __syncthreads();

  if (team >= P.n_teams) return;
  if (P.pdl) {
    // everything above read only the plan; the operands may still be being written by the previous kernel
    pdl_wait();
    if (TMA && item < item_end) stage(item);
  }
  const int slot = tt / LPR, lq = tt % LPR, par = slot & 1;
  const int blk = blk0 + slot;
  const bool active = blk < P.n_blocks;
  // original row of each of the row-block's 4 rows (-1 = none)
  int orow[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int pos = blk * 4 + r;
    orow[r] = !active ? -1 : PERM ? __ldg(P.perm + pos) : (pos < P.n_rows ? pos : -1);
  }
  int my_len = 0;
  if (active) my_len = __ldg(P.blkptr + blk + 1) - __ldg(P.blkptr + blk);
  // uniform trip count of the warp (8 row-blocks), rounded up to the 2-step pipeline
  const int wlen = (__reduce_max_sync(0xffffffffu, my_len) + 1) & ~1;

  // byte offsets of this lane's float4 columns inside a staged row: j = 0/2 through cA, j = 1/3 through cB
  // LPR 4: j = 0/2 through cA, j = 1/3 through cB (XOR parity: the two row-blocks of a quarter-warp phase read
  // opposite 64-byte halves); LPR 8: the 8 lanes of a row-block read one contiguous 128-byte half per column.
  const uint32_t cA = LPR == 4 ? ((uint32_t)(lq * 16) ^ (uint32_t)(par * 64)) : (uint32_t)(lq * 16);
  const uint32_t cB = LPR == 4 ? ((uint32_t)(lq * 16 + 64) ^ (uint32_t)(par * 64)) : (uint32_t)(lq * 16 + 128);
  const uint8_t* xA = xs + cA;
  const uint8_t* xB = xs + cB;
  const uint32_t* po = s_off + slot;
  const float4* pw = s_val + slot;
  // channel of accumulator column j inside the slab
  int ch[NJ];
  ch[0] = (int)(cA >> 2), ch[1] = (int)(cB >> 2);
  if constexpr (NJ == 4) ch[2] = ch[0] + 32, ch[3] = ch[1] + 32;

  uint32_t phase = 0;
  int claim_par = 1;  // slot parity of the claim made during this item (slot 0 held the first one)
  while (item < item_end) {
    // claim the next item now; the index is read after the entry loop, when the round trip is long over
    if (cnt && tt == 0) s_claim[2 * team + claim_par] = atomicAdd(cnt, 1);
    int next_item = item_end;
    const int b = item / P.n_slabs, slab = item - b * P.n_slabs;
    const int slab_f = min(64, a.F - slab * 64);       // channels in this slab (multiple of 4)
    const bool prof = (P.debug_skip == 4) && (tt == 0);
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    if (prof) t0 = clock64();
    if (!TMA) stage(item);
    // Accumulators start at (beta * Z + G) / alpha (alpha is 1 or 2, so the scaling is exact): the
    // loads are issued here, behind the cp.asyncs, and land while the tile is being staged.
    if (prof) t1 = clock64();
    float4 acc[4][NJ];
    {
      const float inv_alpha = 1.f / a.alpha;
      const float zs = a.beta * inv_alpha;
      // No per-element range checks here: rows past the end (or of an inactive row-block) and channels past the slab are
      // CLAMPED to valid addresses — their accumulators are computed on junk and never stored.  (The checked form cost
      // several hundred instructions per item and warp.)
      int chc[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) chc[j] = slab * 64 + min(ch[j], slab_f - 4);
      int64_t rowc[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) rowc[r] = PERM ? max(orow[r], 0) : min(blk * 4 + r, P.n_rows - 1);
      // batch 1: all Z loads in flight together
      if (a.Z != nullptr) {
        const float* zb = a.Z + b * a.z_sB;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float* zr = zb + rowc[r] * a.z_sV;
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const float4 z = ldcg4(zr + chc[j]);
            acc[r][j] = make_float4(z.x * zs, z.y * zs, z.z * zs, z.w * zs);
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int j = 0; j < NJ; ++j) acc[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // batch 2 (adjoint recurrence only): all G loads in flight together
      if (a.G != nullptr) {
        float4 g[4][NJ];
        const float* gb = a.G + b * a.g_sB;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float* gr = gb + rowc[r] * a.g_sV;
#pragma unroll
          for (int j = 0; j < NJ; ++j) g[r][j] = ldcg4(gr + chc[j]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int j = 0; j < NJ; ++j)
            acc[r][j] = make_float4(fmaf(g[r][j].x, inv_alpha, acc[r][j].x), fmaf(g[r][j].y, inv_alpha, acc[r][j].y),
                                    fmaf(g[r][j].z, inv_alpha, acc[r][j].z), fmaf(g[r][j].w, inv_alpha, acc[r][j].w));
      }
    }
    if (TMA && P.debug_skip == 1) {
    } else if (TMA) {
      hop_mbar_wait(hop_smem_u32(s_bar + team), phase);
      phase ^= 1u;
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      team_sync<TEAM_THREADS>(team);
    }

    if (prof) t2 = clock64();
    {
      auto load_x = [&](uint32_t o, float4(&x)[NJ]) {
        x[0] = *reinterpret_cast<const float4*>(xA + o);
        x[1] = *reinterpret_cast<const float4*>(xB + o);
        if constexpr (NJ == 4) {
          x[2] = *reinterpret_cast<const float4*>(xA + o + 128);
          x[3] = *reinterpret_cast<const float4*>(xB + o + 128);
        }
      };
      // software pipeline: offsets two steps ahead, weights / values one step ahead; the panels carry
      // PANEL_PAD zero steps so the over-reads at the tail are harmless.
      float4 x0[NJ], x1[NJ], w0, w1;
      uint32_t o1, o2;
      w0 = pw[0];
      load_x(po[0], x0);
      o1 = po[DSW_TILE_BLOCKS];
#pragma unroll 1
      for (int u = 0; u < (P.debug_skip == 2 ? 0 : wlen); u += 2) {
        w1 = pw[(u + 1) * DSW_TILE_BLOCKS];
        load_x(o1, x1);
        o2 = po[(u + 2) * DSW_TILE_BLOCKS];
        fma_step(acc, w0, x0);
        w0 = pw[(u + 2) * DSW_TILE_BLOCKS];
        load_x(o2, x0);
        o1 = po[(u + 3) * DSW_TILE_BLOCKS];
        fma_step(acc, w1, x1);
      }
    }

    if (TMA) {
      // every lane of the team is done reading the staged rows: the next item's transfer may start
      team_sync<TEAM_THREADS>(team);
      next_item = cnt ? s_claim[2 * team + claim_par] : item + P.n_teams;
      if (next_item < item_end) stage(next_item);
    }
    if (prof) t3 = clock64();
    // ---- epilogue: O = alpha * acc ----
    if (active) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int row = orow[r];
        if (row < 0) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          if (ch[j] >= slab_f) continue;
          const int64_t col = (int64_t)slab * 64 + ch[j];
          float4 o = make_float4(a.alpha * acc[r][j].x, a.alpha * acc[r][j].y, a.alpha * acc[r][j].z, a.alpha * acc[r][j].w);
          if (a.act) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
          *reinterpret_cast<float4*>(a.O + b * a.o_sB + row * a.o_sV + col) = o;
        }
      }
    }
    // every lane of the team is done reading the staged rows before the next item overwrites them
    if (!TMA) {
      team_sync<TEAM_THREADS>(team);
      next_item = cnt ? s_claim[2 * team + claim_par] : item + P.n_teams;
    }
    if (prof) {
      const long long t4 = clock64();
      atomicAdd(&g_hop_prof[0], (unsigned long long)(t1 - t0));  // issue staging
      atomicAdd(&g_hop_prof[1], (unsigned long long)(t2 - t1));  // Z / G loads + wait for the tile
      atomicAdd(&g_hop_prof[2], (unsigned long long)(t3 - t2));  // entry loop
      atomicAdd(&g_hop_prof[3], (unsigned long long)(t4 - t3));  // stores + team barrier
      atomicAdd(&g_hop_prof[4], 1ull);                           // items
    }
    item = next_item;
    claim_par ^= 1;
  }
  if (cnt) {
    // all active teams of the CTA are done claiming (named barrier over the active teams only: the
    // threads of unused teams have exited)
    asm volatile("bar.sync 15, %0;" ::"r"(P.n_teams * TEAM_THREADS) : "memory");
    if (tid == 0) leave();
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static bool vec4_ok(const HopArgs& a) {
  if (a.F % 4) return false;
  if (!aligned16(a.X) || !aligned16(a.O) || (a.x_sB | a.x_sV | a.o_sB | a.o_sV) % 4) return false;
  if (a.Z && (!aligned16(a.Z) || (a.z_sB | a.z_sV) % 4)) return false;
  if (a.G && (!aligned16(a.G) || (a.g_sB | a.g_sV) % 4)) return false;
  if (a.M && (!aligned16(a.M) || (a.m_sB | a.m_sV) % 4)) return false;
  return true;
}

int launch_hop(const dsw_csr& A, const dsw_rb& rb, const HopArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.F <= 0 || !a.X || !a.O) return DSW_ERR_BAD_ARGUMENT;
  if (a.B > 65535) return DSW_ERR_UNSUPPORTED;
  const bool v4 = vec4_ok(a);
  const int hop_mode = (int)g_options[DSW_OPT_HOP_KERNEL].load(std::memory_order_relaxed);
  // 32-bit byte offsets inside one sample: (n_cols - 1) * x_sV * 4 + F * 4 must fit
  const bool off32 = ((int64_t)A.n_cols * a.x_sV + a.F) * 4 < ((int64_t)1 << 32);
  // narrow planes (the 64 -> 2 output layer runs its hops on 4 channels): a 64-channel slab would be
  // mostly padding, the plain CSR kernel moves only the real channels
  const int64_t small_f_opt = g_options[DSW_OPT_HOP_SMALL_F].load(std::memory_order_relaxed);
  const int small_f = small_f_opt > 0 ? (int)small_f_opt : 8;
  // (a masked hop — one per step: the last Clenshaw hop of a ReLU-fed layer's backward — takes the row-block kernel below:
  //  the masked store compiled into the team kernel cost every launch 1-3 %)
  if (v4 && rb.R == 4 && hop_mode == 0 && rb.n_tiles > 0 && a.F > small_f && a.M == nullptr) {
    const size_t panels = (size_t)(rb.tile_len_max + PANEL_PAD) * DSW_TILE_BLOCKS * 20 + (size_t)((rb.tile_rows_max + 1) & ~1) * 4 +
                          (size_t)((rb.tile_pieces_max + 1) & ~1) * 4 + 8 * MAX_TEAMS + 64;
    const size_t xbuf = (size_t)rb.tile_rows_max * 256;
    int n_teams = panels < 200 * 1024 ? (int)std::min<size_t>(MAX_TEAMS, (226 * 1024 - panels) / std::max<size_t>(xbuf, 1)) : 0;
    if (n_teams >= 1) {
      TeamHopPlan P{};
      P.blkptr = rb.blkptr, P.tp_ptr = rb.tp_ptr, P.tp_val = rb.tp_val, P.tp_off = rb.tp_off;
      P.tile_ptr = rb.tile_ptr, P.tile_row = rb.tile_row;
      P.tpc_ptr = rb.tpc_ptr, P.tpc_row = rb.tpc_row, P.tpc_meta = rb.tpc_meta, P.cap_pieces = rb.tile_pieces_max;
      P.n_blocks = rb.n_blocks, P.n_rows = A.n_rows, P.cap_len = rb.tile_len_max, P.cap_rows = rb.tile_rows_max;
      P.n_slabs = ceil_div(a.F, 64);
      P.n_items = a.B * P.n_slabs;
      n_teams = std::min(n_teams, P.n_items);
      if (g_options[DSW_OPT_DEBUG].load(std::memory_order_relaxed) == 3) n_teams = std::min(n_teams, 2);
      {
        const int64_t cap = g_options[DSW_OPT_HOP_TEAMS].load(std::memory_order_relaxed);
        if (cap > 0) n_teams = std::min<int>(n_teams, (int)cap);
      }
      P.n_teams = n_teams;
      P.debug_skip = (int)g_options[DSW_OPT_DEBUG].load(std::memory_order_relaxed);
      P.prefetch_zg = g_options[DSW_OPT_HOP_PREFETCH].load(std::memory_order_relaxed) == 2 ? 0 : 1;
      // items per CTA: a few per team to amortise the staged panels, while keeping >= ~3 waves of CTAs
      int ipc = n_teams;
      while (ipc < 4 * n_teams && ipc * 2 <= P.n_items && (int64_t)rb.n_tiles * ceil_div(P.n_items, ipc * 2) >= 148 * 3)
        ipc *= 2;
      {  // tuning override: items per CTA (rounded up to a multiple of the team count)
        const int64_t o = g_options[DSW_OPT_HOP_IPC].load(std::memory_order_relaxed);
        if (o > 0) ipc = std::min<int>((int)((o + n_teams - 1) / n_teams * n_teams), std::max(P.n_items, n_teams));
      }
      P.items_per_cta = ipc;
      const size_t smem = (size_t)n_teams * xbuf + panels;
      dim3 grid(rb.n_tiles, ceil_div(P.n_items, ipc));
      // Dynamic scheduling (default): the CTAs of a tile claim items from a per-tile counter.  One CTA
      // per tile plus enough extra rows of CTAs to fill the machine and to share the tiles of the last,
      // partial round; CTAs that find their tile finished leave before staging anything.  The counter sets
      // rotate per launch (concurrent streams get different sets); the last CTA of a launch zeroes its set.
      if (rb.hop_cnt && rb.hop_ring && g_options[DSW_OPT_HOP_IPC].load(std::memory_order_relaxed) == 0) {
        const uint32_t k = rb.hop_ring->fetch_add(1, std::memory_order_relaxed);
        P.cnt = rb.hop_cnt + (size_t)(k % HOP_CNT_SLOTS) * (rb.n_tiles + 1);
        const int64_t extra = g_options[DSW_OPT_HOP_ROWS].load(std::memory_order_relaxed);
        const int rows = std::min(ceil_div(P.n_items, n_teams), ceil_div(148, rb.n_tiles) + (extra > 0 ? (int)extra : 2));
        grid = dim3(rb.n_tiles, std::max(rows, 1));
        P.n_ctas = (int32_t)(grid.x * grid.y);
      }
      // tensor maps of the gather source [B][n_cols][F] with boxes of 1 .. 128 rows x 64 channels
      HopMaps maps;
      bool tma = g_options[DSW_OPT_NO_TMA].load(std::memory_order_relaxed) == 0 && rb.tile_pieces_max > 0 &&
                 rb.tile_pieces_max <= rb.tile_rows_max && a.x_sV >= a.F;
      if (tma) {
        const uint64_t dims[3] = {(uint64_t)a.F, (uint64_t)A.n_cols, (uint64_t)a.B};
        const uint64_t strides[2] = {(uint64_t)a.x_sV * 4, (uint64_t)std::max<int64_t>(a.x_sB, 1) * 4};
        for (int j = 0; j < 8 && tma; ++j) {
          const uint32_t box[3] = {64u, 1u << j, 1u};
          tma = (1 << j) > A.n_cols ? (maps.m[j] = maps.m[j - 1], true) : encode_f32_map(&maps.m[j], a.X, 3, dims, strides, box);
        }
      }
      const bool lpr8 = g_options[DSW_OPT_HOP_LPR].load(std::memory_order_relaxed) == 8;
      P.pdl = g_options[DSW_OPT_NO_PDL].load(std::memory_order_relaxed) == 0 ? 1 : 0;
      auto launch = [&](auto kern, int threads, int slot) -> int {
        static PerDeviceOnce attr_done[6];
        DSW_CUDA_TRY(attr_done[slot].max_dynamic_smem(kern, 226 * 1024));
        if (P.pdl) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = grid, cfg.blockDim = dim3(threads), cfg.dynamicSmemBytes = smem, cfg.stream = st;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          attr[0].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = attr, cfg.numAttrs = 1;
          DSW_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, P, a, maps));
        } else {
          kern<<<grid, threads, smem, st>>>(P, a, maps);
        }
        return check_launch();
      };
      P.perm = rb.perm;
      if (rb.perm) {  // permuted plans: 4 lanes per row-block only
        if (tma) return launch(hop_team_kernel<true, 4, true>, DSW_TILE_BLOCKS * 4 * MAX_TEAMS, 4);
        return launch(hop_team_kernel<false, 4, true>, DSW_TILE_BLOCKS * 4 * MAX_TEAMS, 5);
      }
      if (tma) return lpr8 ? launch(hop_team_kernel<true, 8, false>, DSW_TILE_BLOCKS * 8 * MAX_TEAMS, 0)
                           : launch(hop_team_kernel<true, 4, false>, DSW_TILE_BLOCKS * 4 * MAX_TEAMS, 1);
      return lpr8 ? launch(hop_team_kernel<false, 8, false>, DSW_TILE_BLOCKS * 8 * MAX_TEAMS, 2)
                  : launch(hop_team_kernel<false, 4, false>, DSW_TILE_BLOCKS * 4 * MAX_TEAMS, 3);
    }
  }
  if (v4 && rb.R == 4 && hop_mode == 3 && off32 && rb.tile_entries_max > 0 && !rb.perm) {
    const size_t smem = (size_t)((rb.tile_entries_max + 3) & ~3) * 20;
    const int n_tiles = ceil_div(rb.n_blocks, TILE_BLOCKS);
    if (smem <= 200 * 1024) {
      const int nf = a.F > 32 ? 2 : 1;
      const int slabs = ceil_div(a.F, 32 * nf);
      // samples per CTA: amortise the staged panels while keeping >= ~6 CTAs per SM in the grid
      int ns = 1;
      while (ns < 8 && ns * 2 <= a.B && (int64_t)n_tiles * slabs * ceil_div(a.B, ns * 2) >= 148 * 6) ns *= 2;
      dim3 grid(n_tiles, slabs, ceil_div(a.B, ns));
      static PerDeviceOnce attr_set[2];
      if (nf == 2) {
        if (smem > 48 * 1024) DSW_CUDA_TRY(attr_set[1].max_dynamic_smem(hop_tile_kernel<2>, 200 * 1024));
        hop_tile_kernel<2><<<grid, TILE_THREADS, smem, st>>>(rb.blkptr, rb.ucol, reinterpret_cast<const float4*>(rb.uval),
                                                             rb.n_blocks, A.n_rows, ns, a);
      } else {
        if (smem > 48 * 1024) DSW_CUDA_TRY(attr_set[0].max_dynamic_smem(hop_tile_kernel<1>, 200 * 1024));
        hop_tile_kernel<1><<<grid, TILE_THREADS, smem, st>>>(rb.blkptr, rb.ucol, reinterpret_cast<const float4*>(rb.uval),
                                                             rb.n_blocks, A.n_rows, ns, a);
      }
      return check_launch();
    }
  }
  if (v4 && rb.R > 0 && hop_mode <= 1 && a.F > small_f && !rb.perm) {
    const int threads = 512;
    const int slots = threads / 16;
    dim3 grid(ceil_div(rb.n_blocks, slots), ceil_div(a.F, 64), a.B);
    if (rb.R == 4)
      hop_rb_kernel<4><<<grid, threads, 0, st>>>(rb.blkptr, rb.ucol, rb.uval, rb.n_blocks, A.n_rows, a);
    else
      hop_rb_kernel<2><<<grid, threads, 0, st>>>(rb.blkptr, rb.ucol, rb.uval, rb.n_blocks, A.n_rows, a);
    return check_launch();
  }
  if (v4 && a.F == 4 && a.B >= 8) {
    constexpr int NB = 8;
    dim3 grid(ceil_div(A.n_rows, 256), ceil_div(a.B, NB));
    DSW_CUDA_TRY(launch_pdl(hop_csr_f4_kernel<NB>, grid, dim3(256), 0, st, pdl_enabled(), (const int32_t*)A.rowptr,
                            (const int32_t*)A.col, (const float*)A.val, (int32_t)A.n_rows, a));
    return check_launch();
  }
  const int vec = v4 ? 4 : 1;
  const int nvec = a.F / vec;
  int lpr_log2 = 0;
  while ((1 << lpr_log2) < nvec && lpr_log2 < 8) ++lpr_log2;
  const int rows_per_cta = 256 >> lpr_log2;
  dim3 grid(ceil_div(A.n_rows, rows_per_cta), a.B);
  if (v4)
    DSW_CUDA_TRY(launch_pdl(hop_csr_kernel<4>, grid, dim3(256), 0, st, pdl_enabled(), (const int32_t*)A.rowptr, (const int32_t*)A.col,
                            (const float*)A.val, (int32_t)A.n_rows, (int32_t)lpr_log2, a));
  else
    DSW_CUDA_TRY(launch_pdl(hop_csr_kernel<1>, grid, dim3(256), 0, st, pdl_enabled(), (const int32_t*)A.rowptr, (const int32_t*)A.col,
                            (const float*)A.val, (int32_t)A.n_rows, (int32_t)lpr_log2, a));
  return check_launch();
}

}  // namespace dsw

using namespace dsw;

extern "C" {

int dsw_debug_counters(uint64_t* out8, int reset) {
  if (!out8) return DSW_ERR_BAD_ARGUMENT;
  unsigned long long h[8];
  DSW_CUDA_TRY(cudaMemcpyFromSymbol(h, g_hop_prof, sizeof(h)));
  for (int i = 0; i < 8; ++i) out8[i] = h[i];
  if (reset) {
    unsigned long long z[8] = {};
    DSW_CUDA_TRY(cudaMemcpyToSymbol(g_hop_prof, z, sizeof(z)));
  }
  return DSW_OK;
}

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

// Strided-output / accumulating forms (the skip concatenation of the U-Net without a cat: the unpool writes its rows
// straight into one half of the [B, V, 2C] buffer the decoder reads; the pool's backward adds the skip gradient).
int dsw_spmm_fwd_ex(const dsw_plan* mat, const float* x, int64_t x_sB, int64_t x_sV, const float* addend, int64_t a_sB,
                    int64_t a_sV, float* y, int64_t y_sB, int64_t y_sV, int32_t B, int32_t F, void* stream) {
  if (!mat || !x || !y || B <= 0 || F <= 0 || y_sV < F) return DSW_ERR_BAD_ARGUMENT;
  HopArgs a;
  a.X = x, a.x_sB = x_sB, a.x_sV = x_sV;
  a.G = addend, a.g_sB = a_sB, a.g_sV = a_sV;
  a.O = y, a.o_sV = y_sV, a.o_sB = y_sB;
  a.B = B, a.F = F;
  return launch_hop(mat->fwd, mat->fwd_rb, a, static_cast<cudaStream_t>(stream));
}

int dsw_spmm_bwd_ex(const dsw_plan* mat, const float* dy, int64_t dy_sB, int64_t dy_sV, const float* addend, int64_t a_sB,
                    int64_t a_sV, float* dx, int64_t dx_sB, int64_t dx_sV, int32_t B, int32_t F, void* stream) {
  if (!mat || !dy || !dx || B <= 0 || F <= 0 || dx_sV < F) return DSW_ERR_BAD_ARGUMENT;
  HopArgs a;
  a.X = dy, a.x_sB = dy_sB, a.x_sV = dy_sV;
  a.G = addend, a.g_sB = a_sB, a.g_sV = a_sV;
  a.O = dx, a.o_sV = dx_sV, a.o_sB = dx_sB;
  a.B = B, a.F = F;
  return launch_hop(mat->tr, mat->tr_rb, a, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
