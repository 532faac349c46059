// tcgen05 (5th-gen tensor core) channel mix for sm_100a.
//
//   C[n][c] = bias[c] + sum_p sum_kk A_p[n][kk] * B[p][kk][c]         (MixArgs, dsw_internal.cuh)
//
// the dense (K*Fin)->Fout contraction of conv_cheb (reference modules/layers.py:171-177) and its
// transpose dy.W^T for the input gradient.  fp32 accuracy on bf16 tensor cores by operand splitting:
//   a = a_hi + a_lo,  b = b_hi + b_lo  (bf16 each);   a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi
// three tcgen05.mma per K-step accumulate into one fp32 TMEM tile; the dropped a_lo*b_lo term and the
// second-order rounding leave a relative error of a few 1e-6 per product (parity bar: 1e-4).
//
// Two kernels (see the comments above each): mix_tma_kernel, the production path — persistent,
// warp-specialised, the fp32 A operand landed by tensor-map TMA and converted to bf16 hi / lo in place
// — and mix_tc_kernel, the same pipeline with register-path converters for operands that TMA cannot
// describe (unaligned pointers / strides).  B (the weights) is split and laid out as ready-to-copy
// shared-memory images by a small prep kernel once per call.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>

#include "dsw_internal.cuh"
#include "dsw_tmap.cuh"

namespace dsw {
namespace tc {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BKB = 64;          // reduction elements per stage: 64 bf16 = one 128-byte swizzle row
constexpr int A_TILE = BM * 128; // bytes of one bf16 A image (hi or lo)

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 128-byte rows,
// 8-row core groups 1024 bytes apart (SBO), descriptor version 1 (Blackwell), layout type 2.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;      // stride byte offset
  d |= (uint64_t)1 << 46;                // version
  d |= (uint64_t)2 << 61;                // SWIZZLE_128B
  return d;
}

// byte offset of the 16-byte chunk `chunk` (0..7) of row `row` inside a swizzled [rows][128 B] image
__device__ __host__ __forceinline__ uint32_t swz(uint32_t row, uint32_t chunk) {
  return row * 128u + (((chunk ^ (row & 7u)) & 7u) << 4);
}

// streaming 16-byte load that does not allocate an L1 line (the A operand is read exactly once)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

struct TcArgs {
  MixArgs m;
  const uint8_t* bprep;  // [n_tiles][P * nkb][hi image | lo image], image = BN x 128 B swizzled
  int32_t BN;            // columns per CTA (multiple of 16, <= 256)
  int32_t nkb;           // 64-wide reduction blocks per plane
  int32_t tmem_cols;     // power of two >= max(32, BN); two accumulators are allocated
  int32_t stages;        // shared-memory ring depth
  int32_t n_ctiles;      // column tiles
  int32_t dbg;           // timing experiments only (DSW_OPT_DEBUG bits 16..256; results become wrong)
  int32_t cmode;         // split_pair mode
};

// ---------------------------------------------------------------------------------------------
// B preparation: split weights into bf16 hi/lo and write shared-memory images
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mix_tc_prep_kernel(TcArgs P, uint8_t* __restrict__ out, int32_t n_tiles) {
  pdl_trigger();
  pdl_wait();
  const MixArgs& a = P.m;
  const int64_t chunks_per_img = (int64_t)P.BN * 8;  // 16-byte chunks
  const int64_t total = (int64_t)n_tiles * a.P * P.nkb * chunks_per_img;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int chunk = (int)(i % 8);
  const int row = (int)((i / 8) % P.BN);
  const int64_t blk = i / chunks_per_img;  // (tile, p, kb)
  const int kb = (int)(blk % P.nkb);
  const int p = (int)((blk / P.nkb) % a.P);
  const int tile = (int)(blk / ((int64_t)P.nkb * a.P));
  const int c = tile * P.BN + row;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat16 h[2], l[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int kk = kb * BKB + chunk * 8 + j * 2 + q;
      float v = 0.f;
      if (kk < a.Ka && c < a.Nc) {
        const int cp = c / a.Cw, cc = c - cp * a.Cw;
        v = __ldg(a.Bm + p * a.sBp + (int64_t)kk * a.sBk + cp * a.sBc1 + cc * a.sBc0);
      }
      split_bf16(v, h[q], l[q]);
    }
    hi[j] = pack2(h[0], h[1]);
    lo[j] = pack2(l[0], l[1]);
  }
  const int64_t img_bytes = (int64_t)P.BN * 128;
  uint8_t* base = out + blk * 2 * img_bytes;
  *reinterpret_cast<uint4*>(base + swz(row, chunk)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(base + img_bytes + swz(row, chunk)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ---------------------------------------------------------------------------------------------
// main kernel: persistent, warp-specialised
//
//   warps 0-7   converters   fp32 A rows (global, coalesced float4, one 64-wide reduction block kept in
//                            flight in registers) -> bf16 hi / lo -> shared memory stage (UMMA K-major
//                            SWIZZLE_128B); arrive on a_full[stage]
//   warp  8     B loader     one lane: cp.async.bulk of the pre-split weight block [hi | lo] into the
//                            stage, completion (tx bytes) on b_full[stage]
//   warp  9     MMA issuer   one lane: 3 x tcgen05.mma per 16-wide K step into the TMEM accumulator of
//                            the current tile (two accumulators, ping-pong); tcgen05.commit releases the
//                            stage (empty[stage]) and, after a tile's last block, hands the accumulator
//                            to the epilogue (acc_full[buf])
//   warps 10-13 epilogue     tcgen05.ld -> + bias / ReLU -> global; arrive on acc_empty[buf]
//
// CTAs are persistent (grid = #SMs) and walk the (row tile, column tile) list with a static stride, so
// barrier / TMEM set-up happens once and the converters stream straight across tile boundaries while
// the epilogue of the previous tile drains the other accumulator.
// ---------------------------------------------------------------------------------------------
constexpr int N_CONV_WARPS = 8;
constexpr int WARP_BLOAD = 8, WARP_MMA = 9;  // warps 10-13: epilogue
constexpr int THREADS2 = 32 * 14;

template <bool VEC4>
__global__ void __launch_bounds__(THREADS2, 1) mix_tc_kernel(const __grid_constant__ TcArgs P) {
  extern __shared__ uint8_t smem_raw[];
  const MixArgs& a = P.m;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int BN = P.BN;
  const int S = P.stages;
  const int cmode = P.cmode;
  const uint32_t b_img = (uint32_t)BN * 128u;
  const uint32_t stage_bytes = 2u * A_TILE + 2u * b_img;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + (uint32_t)S * stage_bytes;
  // barrier slots (8 bytes each): a_full[S], b_full[S], empty[S], acc_full[2], acc_empty[2]; then the TMEM slot
  auto a_full = [&](int s) { return bars + 8u * s; };
  auto b_full = [&](int s) { return bars + 8u * (S + s); };
  auto empty = [&](int s) { return bars + 8u * (2 * S + s); };
  auto acc_full = [&](int b) { return bars + 8u * (3 * S + b); };
  auto acc_empty = [&](int b) { return bars + 8u * (3 * S + 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (3 * S + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (size_t)S * stage_bytes + 8 * (3 * S + 4));

  if (t == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(a_full(s), N_CONV_WARPS * 32);
      mbar_init(b_full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 128);
    }
    fence_mbar_init();
  }
  if (warp == WARP_MMA) tmem_alloc(tmem_slot, (uint32_t)(2 * P.tmem_cols));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int total_kb = a.P * P.nkb;
  const int last_ksteps = (a.Ka - (P.nkb - 1) * BKB + 15) / 16;  // K-steps (of 16) in a plane's last block
  const int n_ctiles = P.n_ctiles;
  const int64_t n_rtiles = (a.N + BM - 1) / BM;
  const int64_t n_tiles = n_rtiles * n_ctiles;

  if (warp < N_CONV_WARPS) {
    // ================= converters =================
    // thread -> rows (t>>4) + 16*i, i = 0..7; float4 column q = t & 15 (reduction elements 4q..4q+3)
    const int q = t & 15;
    float4 cur[8];
    int32_t rb[8], rv[8];
    bool rok[8];
    auto set_rows = [&](int64_t tile) {
      const int64_t n0 = (tile / n_ctiles) * BM;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t n = n0 + (t >> 4) + 16 * i;
        rok[i] = n < a.N;
        const int64_t bb = rok[i] ? n / a.rows_per_batch : 0;
        rb[i] = (int32_t)bb;
        rv[i] = rok[i] ? (int32_t)(n - bb * a.rows_per_batch) : 0;
      }
    };
    auto load_block = [&](int kbi, float4(&dst)[8]) {
      const int p = kbi / P.nkb, kb = kbi - p * P.nkb;
      const int kk = kb * BKB + q * 4;
      const float* __restrict__ Ap = a.A[p];
      const int64_t sB = a.a_sB[p], sV = a.a_sV[p];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rok[i]) {
          const float* src = Ap + rb[i] * sB + rv[i] * sV + kk;
          if (VEC4) {
            if (kk < a.Ka) v = ld_stream4(src);  // Ka % 4 == 0; streamed once
          } else {
            if (kk + 0 < a.Ka) v.x = __ldg(src + 0);
            if (kk + 1 < a.Ka) v.y = __ldg(src + 1);
            if (kk + 2 < a.Ka) v.z = __ldg(src + 2);
            if (kk + 3 < a.Ka) v.w = __ldg(src + 3);
          }
        }
        dst[i] = v;
      }
    };
    // Two 64-wide blocks are kept in flight per thread (64 KB per SM): (tile, block) pairs are walked
    // as one flat sequence, so the prefetch runs straight across tile boundaries.
    float4 nxt[8];
    int64_t ld_tile = blockIdx.x;  // position of the next block to *load*
    int ld_kbi = 0;
    auto advance_load = [&](float4(&dst)[8]) {
      if (ld_tile >= n_tiles) return;
      if (ld_kbi == 0) set_rows(ld_tile);
      load_block(ld_kbi, dst);
      if (++ld_kbi == total_kb) ld_kbi = 0, ld_tile += gridDim.x;
    };
    advance_load(cur);
    advance_load(nxt);
    auto convert_store = [&](int s, const float4(&src)[8]) {
      uint8_t* Ahi = smem_gen + (size_t)s * stage_bytes;
      uint8_t* Alo = Ahi + A_TILE;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t row = (t >> 4) + 16 * i;
        uint2 qh, ql;
        split_quad(src[i], qh, ql, cmode);
        const uint32_t off = swz(row, q >> 1) + (q & 1) * 8;
        *reinterpret_cast<uint2*>(Ahi + off) = qh;
        *reinterpret_cast<uint2*>(Alo + off) = ql;
      }
    };
    const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int64_t n_iters = my_tiles * total_kb;
    for (int64_t it2 = 0; it2 < n_iters; it2 += 2) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int64_t itx = it2 + half;
        if (itx >= n_iters) break;
        const int s = (int)(itx % S);
        const uint32_t use = (uint32_t)(itx / S);  // how many times this stage has been filled before
        if (use > 0) {
          mbar_wait(empty(s), (use - 1) & 1);
          tc_fence_after();
        }
        if (half == 0) {
          convert_store(s, cur);
          advance_load(cur);
        } else {
          convert_store(s, nxt);
          advance_load(nxt);
        }
        fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
        mbar_arrive(a_full(s));
      }
    }
  } else if (warp == WARP_BLOAD) {
    // ================= B loader (one thread) =================
    if (lane == 0) {
      int64_t it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int ctile = (int)(tile % n_ctiles);
        const uint8_t* bsrc = P.bprep + (int64_t)ctile * total_kb * 2 * b_img;
        for (int kbi = 0; kbi < total_kb; ++kbi, ++it) {
          const int s = (int)(it % S);
          const uint32_t use = (uint32_t)(it / S);
          if (use > 0) mbar_wait(empty(s), (use - 1) & 1);
          mbar_expect_tx(b_full(s), 2u * b_img);
          bulk_copy_g2s(smem_base + s * stage_bytes + 2u * A_TILE, bsrc + (int64_t)kbi * 2 * b_img, 2u * b_img, b_full(s));
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int64_t it = 0;
      int64_t local_tile = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++local_tile) {
        const int buf = (int)(local_tile & 1);
        const uint32_t buse = (uint32_t)(local_tile >> 1);
        if (buse > 0) {
          mbar_wait(acc_empty(buf), (buse - 1) & 1);
          tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * P.tmem_cols);
        for (int kbi = 0; kbi < total_kb; ++kbi, ++it) {
          const int s = (int)(it % S);
          const uint32_t ph = (uint32_t)(it / S) & 1;
          mbar_wait(a_full(s), ph);
          mbar_wait(b_full(s), ph);
          tc_fence_after();
          const uint32_t st_base = smem_base + s * stage_bytes;
          const int kb = kbi % P.nkb;
          const int ksteps = (kb == P.nkb - 1) ? last_ksteps : BKB / 16;
          const uint64_t dAh = make_desc(st_base), dAl = make_desc(st_base + A_TILE);
          const uint64_t dBh = make_desc(st_base + 2u * A_TILE), dBl = make_desc(st_base + 2u * A_TILE + b_img);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t adv = (uint64_t)(ks * 2);  // 32 bytes per K-step, in 16-byte units
            umma_bf16(d_tmem, dAh + adv, dBh + adv, idesc, (kbi | ks) != 0);
            umma_bf16(d_tmem, dAh + adv, dBl + adv, idesc, 1u);
            umma_bf16(d_tmem, dAl + adv, dBh + adv, idesc, 1u);
          }
          umma_commit(empty(s));  // arrives when every MMA issued so far has completed
        }
        umma_commit(acc_full(buf));
      }
    }
  } else {
    // ================= epilogue (4 warps = the 4 TMEM lane quarters) =================
    const int quarter = warp & 3;  // tcgen05.ld: a warp reads lanes 32*(warp%4) .. +31
    const bool vec_store = (a.Cw % 4 == 0) && (a.ldc % 4 == 0) && (a.sCp % 4 == 0) &&
                           ((reinterpret_cast<uintptr_t>(a.C) & 15) == 0);
    const int chunks = BN / 16;
    int64_t local_tile = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++local_tile) {
      const int buf = (int)(local_tile & 1);
      const uint32_t ph = (uint32_t)(local_tile >> 1) & 1;
      const int ctile = (int)(tile % n_ctiles);
      const int64_t n = (tile / n_ctiles) * BM + quarter * 32 + lane;  // output row of this thread
      mbar_wait(acc_full(buf), ph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)(buf * P.tmem_cols) + ((uint32_t)(quarter * 32) << 16);
      for (int ch = 0; ch < chunks; ++ch) {
        uint32_t r[16];
        tmem_ld16(t_addr + (uint32_t)(ch * 16), r);
        tmem_ld_wait();
        if (n >= a.N) continue;
        const int cg0 = ctile * BN + ch * 16;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const int cg = cg0 + j;
          if (cg >= a.Nc) break;
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[e] = __uint_as_float(r[j + e]);
            if (a.bias && cg + e < a.bias_n) v[e] += __ldg(a.bias + cg + e);
            if (a.R && cg + e < a.Nc) {
              const float rr = __ldg(a.R + n * a.ldr + cg + e);
              v[e] = a.r_mode == 1 ? (rr > 0.f ? v[e] : 0.f) : fmaf(a.r_scale ? __ldg(a.r_scale) : 1.f, rr, v[e]);
            }
            if (a.act == 1) v[e] = fmaxf(v[e], 0.f);
          }
          if (vec_store && cg + 3 < a.Nc) {
            const int cp = cg / a.Cw, cc = cg - cp * a.Cw;
            *reinterpret_cast<float4*>(a.C + cp * a.sCp + n * a.ldc + cc) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (cg + e >= a.Nc) break;
              const int cp = (cg + e) / a.Cw, cc = (cg + e) - cp * a.Cw;
              a.C[cp * a.sCp + n * a.ldc + cc] = v[e];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty(buf));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) tmem_dealloc(tmem_base, (uint32_t)(2 * P.tmem_cols));
}

// ---------------------------------------------------------------------------------------------
// TMA-fed variant (the production path when the A planes are 16-byte aligned with regular strides).
//
// The fp32 A block [128 rows x 64 reduction elements] is no longer pulled through registers / L1 by
// the converters: one producer lane issues a tensor-map TMA load (cp.async.bulk.tensor, zero-filled
// out of bounds) that lands the raw block in the stage buffer, S stages (up to 128 KB per SM) ahead
// of the arithmetic.  The converters read the raw block from shared memory, synchronise among
// themselves, and overwrite the same 32 KB in place with the bf16 hi / lo UMMA images (a raw fp32
// block and its two bf16 images have the same size).  The epilogue transposes each 32 x 64
// accumulator chunk through shared memory so that global stores are full 256-byte row segments.
//
//   warps 0-7 converters | warp 8 producer (TMA A + bulk B) | warp 9 MMA issuer | warps 10-13 epilogue
// ---------------------------------------------------------------------------------------------
struct TmaArgs {
  TcArgs tc;
  int32_t rank;  // 2: rows = flat (b, v) index; 3: (k, v, b) coordinates
  CUtensorMap amap[DSW_MAX_K];
};

constexpr int EPI_PITCH = 64;                          // floats per staged row; 16-byte chunks XOR-swizzled by the row
constexpr int EPI_BYTES = 32 * EPI_PITCH * 4;          // one warp's 32 x 64 staging chunk
__device__ __forceinline__ int epi_off(int row, int chunk) { return row * EPI_PITCH + ((chunk ^ (row & 15)) << 2); }

// Role timing for tuning (DSW_OPT_DEBUG bit 2048): cycles summed over CTAs, read by dsw_debug_mix_counters:
// [0] producer waiting for a free stage, [1] converter warp 0 waiting for the TMA, [2] converting, [3] MMA issuer waiting
// for a free accumulator, [4] MMA issuer waiting for operands, [5] epilogue warp 0 waiting for the accumulator,
// [6] its TMEM -> shared-memory phase, [7] its store phase, [8] tiles, [9] kernel cycles (CTA 0).
__device__ unsigned long long g_mix_prof[16];

// ADDEND: the epilogue also adds r_scale * R (MixArgs); a separate instantiation so that the plain kernel's
// register allocation is untouched.
template <bool ADDEND>
__global__ void __launch_bounds__(THREADS2, 1) mix_tma_kernel(const __grid_constant__ TmaArgs Q) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  const TcArgs& P = Q.tc;
  const MixArgs& a = P.m;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int BN = P.BN;
  const int S = P.stages;
  const int cmode = P.cmode;
  const uint32_t b_img = (uint32_t)BN * 128u;
  const uint32_t stage_bytes = 2u * A_TILE + 2u * b_img;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* epi_gen = smem_gen + (size_t)S * stage_bytes;
  const uint32_t bars = smem_base + (uint32_t)S * stage_bytes + 4u * EPI_BYTES;
  // barrier slots (8 bytes each): raw_full[S], a_full[S], b_full[S], empty[S], acc_full[2], acc_empty[2]; TMEM slot
  auto raw_full = [&](int s) { return bars + 8u * s; };
  auto a_full = [&](int s) { return bars + 8u * (S + s); };
  auto b_full = [&](int s) { return bars + 8u * (2 * S + s); };
  auto empty = [&](int s) { return bars + 8u * (3 * S + s); };
  auto acc_full = [&](int b) { return bars + 8u * (4 * S + b); };
  auto acc_empty = [&](int b) { return bars + 8u * (4 * S + 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (4 * S + 4);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (size_t)S * stage_bytes + 4 * EPI_BYTES + 8 * (4 * S + 4));

  if (t == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(raw_full(s), 1);
      mbar_init(a_full(s), N_CONV_WARPS);  // one elected arrival per converter warp
      mbar_init(b_full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 4);  // one elected arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == WARP_MMA) tmem_alloc(tmem_slot, (uint32_t)(2 * P.tmem_cols));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();  // barriers and TMEM are set up; the operands may still be being written by the previous kernel
  const bool prof = (P.dbg & 2048) != 0;
  const long long t_start = prof ? clock64() : 0;

  const int total_kb = a.P * P.nkb;
  const int last_ksteps = (a.Ka - (P.nkb - 1) * BKB + 15) / 16;  // K-steps (of 16) in a plane's last block
  const int n_ctiles = P.n_ctiles;
  const int64_t n_rtiles = (a.N + BM - 1) / BM;
  const int64_t n_tiles = n_rtiles * n_ctiles;
  const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t n_iters = my_tiles * total_kb;

  if (warp < N_CONV_WARPS) {
    // ================= converters: raw fp32 block (shared) -> bf16 hi / lo images, in place =================
    const int q = t & 15;
    int s = -1;
    uint32_t ph = 1;  // (stage, phase) advance without 64-bit divisions
    for (int64_t it = 0; it < n_iters; ++it) {
      if (++s == S) s = 0;
      if (s == 0) ph ^= 1;
      uint8_t* Ahi = smem_gen + (size_t)s * stage_bytes;
      uint8_t* Alo = Ahi + A_TILE;
      const long long tc0 = (prof && t == 0) ? clock64() : 0;
      mbar_wait(raw_full(s), ph);
      const long long tc1 = (prof && t == 0) ? clock64() : 0;
      if (prof && t == 0) atomicAdd(&g_mix_prof[1], (unsigned long long)(tc1 - tc0));
      if (P.dbg & 128) {  // timing experiment: no conversion
        if (lane == 0) mbar_arrive(a_full(s));
        continue;
      }
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(Ahi + ((t >> 4) + 16 * i) * 256 + q * 16);
      asm volatile("bar.sync 1, %0;" ::"n"(N_CONV_WARPS * 32) : "memory");  // every converter has read its share
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t row = (t >> 4) + 16 * i;
        uint2 qh, ql;
        split_quad(v[i], qh, ql, cmode);
        const uint32_t off = swz(row, q >> 1) + (q & 1) * 8;
        *reinterpret_cast<uint2*>(Ahi + off) = qh;
        *reinterpret_cast<uint2*>(Alo + off) = ql;
      }
      fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(s));
      if (prof && t == 0) atomicAdd(&g_mix_prof[2], (unsigned long long)(clock64() - tc1));
    }
  } else if (warp == WARP_BLOAD) {
    // ================= producer (one thread): TMA of the raw A block + bulk copy of the B images =================
    if (lane == 0) {
      int s = -1;
      uint32_t use = 0xffffffffu;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t rtile = n_ctiles == 1 ? tile : tile / n_ctiles;
        const int ctile = n_ctiles == 1 ? 0 : (int)(tile - rtile * n_ctiles);
        const int64_t n0 = rtile * BM;
        const uint8_t* bsrc = P.bprep + (int64_t)ctile * total_kb * 2 * b_img;
        for (int kbi = 0; kbi < total_kb; ++kbi) {
          if (++s == S) s = 0;
          if (s == 0) ++use;
          if (use > 0) {
            const long long tq = prof ? clock64() : 0;
            mbar_wait(empty(s), (use - 1) & 1);
            if (prof) atomicAdd(&g_mix_prof[0], (unsigned long long)(clock64() - tq));
          }
          const int p = kbi / P.nkb, kb = kbi - p * P.nkb;
          const uint32_t st_base = smem_base + s * stage_bytes;
          if (P.dbg & 32) {  // timing experiment: no A transfer
            mbar_arrive(raw_full(s));
          } else {
            mbar_expect_tx(raw_full(s), 2u * A_TILE);
            if (Q.rank == 2) {
              tma_load_2d(st_base, &Q.amap[p], kb * BKB, (int)n0, raw_full(s));
            } else {
              const int bb = (int)(n0 / a.rows_per_batch);
              tma_load_3d(st_base, &Q.amap[p], kb * BKB, (int)(n0 - (int64_t)bb * a.rows_per_batch), bb, raw_full(s));
            }
          }
          if (P.dbg & 64) {  // timing experiment: no B transfer
            mbar_arrive(b_full(s));
          } else {
            mbar_expect_tx(b_full(s), 2u * b_img);
            bulk_copy_g2s(st_base + 2u * A_TILE, bsrc + (int64_t)kbi * 2 * b_img, 2u * b_img, b_full(s));
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int s = -1;
      uint32_t ph = 1;
      for (int64_t lt = 0; lt < my_tiles; ++lt) {
        const int buf = (int)(lt & 1);
        const uint32_t buse = (uint32_t)(lt >> 1);
        if (buse > 0) {
          const long long tm = prof ? clock64() : 0;
          mbar_wait(acc_empty(buf), (buse - 1) & 1);
          tc_fence_after();
          if (prof) atomicAdd(&g_mix_prof[3], (unsigned long long)(clock64() - tm));
        }
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * P.tmem_cols);
        for (int kbi = 0; kbi < total_kb; ++kbi) {
          if (++s == S) s = 0;
          if (s == 0) ph ^= 1;
          const long long tm = prof ? clock64() : 0;
          mbar_wait(a_full(s), ph);
          mbar_wait(b_full(s), ph);
          tc_fence_after();
          if (prof) atomicAdd(&g_mix_prof[4], (unsigned long long)(clock64() - tm));
          const uint32_t st_base = smem_base + s * stage_bytes;
          const int kb = kbi % P.nkb;
          const int ksteps = (kb == P.nkb - 1) ? last_ksteps : BKB / 16;
          const uint64_t dAh = make_desc(st_base), dAl = make_desc(st_base + A_TILE);
          const uint64_t dBh = make_desc(st_base + 2u * A_TILE), dBl = make_desc(st_base + 2u * A_TILE + b_img);
          for (int ks = 0; ks < ((P.dbg & 256) ? 0 : ksteps); ++ks) {
            const uint64_t adv = (uint64_t)(ks * 2);  // 32 bytes per K-step, in 16-byte units
            umma_bf16(d_tmem, dAh + adv, dBh + adv, idesc, (kbi | ks) != 0);
            umma_bf16(d_tmem, dAh + adv, dBl + adv, idesc, 1u);
            umma_bf16(d_tmem, dAl + adv, dBh + adv, idesc, 1u);
          }
          umma_commit(empty(s));  // arrives when every MMA issued so far has completed
        }
        umma_commit(acc_full(buf));
      }
    }
  } else {
    // ================= epilogue (4 warps = the 4 TMEM lane quarters) =================
    const int quarter = warp & 3;  // tcgen05.ld: a warp reads lanes 32*(warp%4) .. +31
    float* stg = reinterpret_cast<float*>(epi_gen + (size_t)quarter * EPI_BYTES);
    const bool vec_ok = (a.Cw % 4 == 0) && (a.ldc % 4 == 0) && (a.sCp % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(a.C) & 15) == 0);
    const int half = lane >> 4, c4 = lane & 15;  // store phase: two rows per instruction, 16 lanes x float4 per row
    const bool single_plane = a.Cw >= a.Nc;
    // ADDEND: the epilogue is bound by the latency of its addend loads (short-K mixes: 8 KB in flight per SM).  Each warp
    // pulls the addend rows of its NEXT tile into L2 a whole tile ahead (one row per lane, a 128-byte line per prefetch).
    auto prefetch_addend = [&](int64_t lt_next) {
      if (!ADDEND || a.R == nullptr || lt_next >= my_tiles || (P.dbg & 1024)) return;
      const int64_t tile_n = blockIdx.x + lt_next * gridDim.x;
      const int64_t rt_n = n_ctiles == 1 ? tile_n : tile_n / n_ctiles;
      const int64_t n = rt_n * BM + quarter * 32 + lane;
      const int c0 = (int)(tile_n - rt_n * n_ctiles) * BN;
      if (n >= a.N) return;
      const float* p = a.R + n * a.ldr + c0;
      const int ncols = min(BN, a.Nc - c0);
      for (int c = 0; c < ncols; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + c));
    };
    prefetch_addend(0);
    const float rs_all = (ADDEND && a.r_scale != nullptr) ? __ldg(a.r_scale) : 1.f;
    for (int64_t lt = 0; lt < my_tiles; ++lt) {
      const int64_t tile = blockIdx.x + lt * gridDim.x;
      const int buf = (int)(lt & 1);
      const uint32_t ph = (uint32_t)(lt >> 1) & 1;
      const int64_t rtile = n_ctiles == 1 ? tile : tile / n_ctiles;
      const int ctile = n_ctiles == 1 ? 0 : (int)(tile - rtile * n_ctiles);
      const int64_t row0 = rtile * BM + quarter * 32;  // first output row of this warp
      prefetch_addend(lt + 1);
      const bool ep = prof && quarter == 0 && lane == 0;
      long long te0 = ep ? clock64() : 0, te_ld = 0, te_st = 0;
      mbar_wait(acc_full(buf), ph);
      tc_fence_after();
      if (ep) atomicAdd(&g_mix_prof[5], (unsigned long long)(clock64() - te0)), atomicAdd(&g_mix_prof[8], 1ull);
      const uint32_t t_addr = tmem_base + (uint32_t)(buf * P.tmem_cols) + ((uint32_t)(quarter * 32) << 16);
      for (int cg = 0; cg < BN; cg += 64) {
        const int ncol = min(64, BN - cg);  // multiple of 16 (warp-uniform)
        // columns of this lane in the store phase
        const int col = ctile * BN + cg + c4 * 4;
        const bool col_ok = (c4 * 4 < ncol) && (col < a.Nc);
        // the bias values of this lane's columns: requested before the accumulator chunk is fetched, so that their
        // latency (a dependent global load per chunk: ~500 cycles, a third of a short-K tile's epilogue) is hidden
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (col_ok && a.bias != nullptr) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (col + e < a.bias_n) bv[e] = __ldg(a.bias + col + e);
        }
        // ADDEND: the addend row segments of this lane (16 rows x 16 bytes) are requested in two batches of 8: the first
        // before the accumulator chunk moves TMEM -> shared memory, the second one by one as the first is consumed.  Only
        // whole 32-row blocks take this path (the last, partial tile falls through to the scalar tail): the epilogue
        // warps run alone on their schedulers, so the loop is bound by its instruction count and register pressure.
        float4 rv[8];
        bool addend_vec = false;
        const float* r_row = nullptr;
        if (ADDEND) {
          const int ccol0 = single_plane ? col : col % a.Cw;
          addend_vec = col_ok && a.R != nullptr && vec_ok && (col + 3 < a.Nc) && (ccol0 + 3 < a.Cw) && (a.ldr % 4 == 0) &&
                       ((reinterpret_cast<uintptr_t>(a.R) & 15) == 0) && (row0 + 32 <= a.N) && a.act == 0;
          if (addend_vec) {
            r_row = a.R + (row0 + half) * a.ldr + col;
#pragma unroll
            for (int i = 0; i < 8; ++i) rv[i] = __ldg(reinterpret_cast<const float4*>(r_row + (int64_t)(2 * i) * a.ldr));
          }
        }
        // TMEM loads two chunks (32 columns) at a time, one wait per pair
        if (ep) te0 = clock64();
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {
          if (hq * 32 >= ncol) break;
          uint32_t r[2][16];
          tmem_ld16(t_addr + (uint32_t)(cg + hq * 32), r[0]);
          if (hq * 32 + 16 < ncol) tmem_ld16(t_addr + (uint32_t)(cg + hq * 32 + 16), r[1]);
          tmem_ld_wait();
#pragma unroll
          for (int q2 = 0; q2 < 2; ++q2)
            if (hq * 32 + q2 * 16 < ncol) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<uint4*>(stg + epi_off(lane, (hq * 32 + q2 * 16 + j) >> 2)) =
                    make_uint4(r[q2][j], r[q2][j + 1], r[q2][j + 2], r[q2][j + 3]);
            }
        }
        __syncwarp();
        if (ep) te_ld += clock64() - te0, te0 = clock64();
        if (col_ok) {
          const int cp = single_plane ? 0 : col / a.Cw, ccol = col - cp * a.Cw;
          const bool vec = vec_ok && (col + 3 < a.Nc) && (ccol + 3 < a.Cw);
          float* cbase = a.C + (int64_t)cp * a.sCp + ccol;
          const bool relu = a.act == 1;
          if (ADDEND && addend_vec) {
            // same, plus the addend row segments (C += r_scale * R, or the ReLU mask C = R > 0 ? C : 0)
            const float rs = rs_all;
            const bool mask_mode = a.r_mode == 1;
            float* cp_row = cbase + (row0 + half) * a.ldc;
            const int64_t step = 2 * (int64_t)a.ldc;
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              float4 v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(stg + epi_off(2 * (h8 * 8 + i) + half, c4));
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 r4 = rv[i];
                if (h8 == 0) rv[i] = __ldg(reinterpret_cast<const float4*>(r_row + (int64_t)(2 * (8 + i)) * a.ldr));
                float4 o;
                if (mask_mode) {
                  o = make_float4(r4.x > 0.f ? v[i].x + bv[0] : 0.f, r4.y > 0.f ? v[i].y + bv[1] : 0.f,
                                  r4.z > 0.f ? v[i].z + bv[2] : 0.f, r4.w > 0.f ? v[i].w + bv[3] : 0.f);
                } else {
                  o = make_float4(fmaf(rs, r4.x, v[i].x + bv[0]), fmaf(rs, r4.y, v[i].y + bv[1]),
                                  fmaf(rs, r4.z, v[i].z + bv[2]), fmaf(rs, r4.w, v[i].w + bv[3]));
                }
                if (!(P.dbg & 16)) *reinterpret_cast<float4*>(cp_row + (int64_t)(h8 * 8 + i) * step) = o;
              }
            }
          } else if (vec && !(ADDEND && a.R != nullptr)) {
            // two phases per 16 rows so that 8 shared-memory reads, then 8 row-segment stores, overlap
            float* cp_row = cbase + (row0 + half) * a.ldc;
            const int64_t step = 2 * (int64_t)a.ldc;
            int64_t n = row0 + half;
            if (row0 + 32 <= a.N && !(P.dbg & 16)) {
              // whole 32-row block in range (every tile but the last): no per-row checks.  The epilogue warps run alone on
              // their schedulers, so this loop is bound by its instruction count (measured: 325 -> ~150 per chunk).
#pragma unroll
              for (int i0 = 0; i0 < 16; i0 += 8) {
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(stg + epi_off(2 * (i0 + i) + half, c4));
                if (relu) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float4 o = make_float4(fmaxf(v[i].x + bv[0], 0.f), fmaxf(v[i].y + bv[1], 0.f),
                                                 fmaxf(v[i].z + bv[2], 0.f), fmaxf(v[i].w + bv[3], 0.f));
                    *reinterpret_cast<float4*>(cp_row + (int64_t)(i0 + i) * step) = o;
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float4 o = make_float4(v[i].x + bv[0], v[i].y + bv[1], v[i].z + bv[2], v[i].w + bv[3]);
                    *reinterpret_cast<float4*>(cp_row + (int64_t)(i0 + i) * step) = o;
                  }
                }
              }
            } else
#pragma unroll
            for (int i0 = 0; i0 < 16; i0 += 8) {
              float4 v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(stg + epi_off(2 * (i0 + i) + half, c4));
#pragma unroll
              for (int i = 0; i < 8; ++i, n += 2, cp_row += step) {
                if (n >= a.N || (P.dbg & 16)) continue;  // dbg 16: timing experiment without the output stores
                float4 o = v[i];
                o.x += bv[0], o.y += bv[1], o.z += bv[2], o.w += bv[3];
                if (relu) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
                *reinterpret_cast<float4*>(cp_row) = o;
              }
            }
          } else {
#pragma unroll 1
            for (int rr = 0; rr < 32; rr += 2) {
              const int64_t n = row0 + rr + half;
              if (n >= a.N) continue;
              const float4 v = *reinterpret_cast<const float4*>(stg + epi_off(rr + half, c4));
              float ve[4] = {v.x + bv[0], v.y + bv[1], v.z + bv[2], v.w + bv[3]};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (col + e >= a.Nc) break;
                if (ADDEND && a.R) {
                  const float rr = __ldg(a.R + n * a.ldr + col + e);
                  ve[e] = a.r_mode == 1 ? (rr > 0.f ? ve[e] : 0.f) : fmaf(a.r_scale ? __ldg(a.r_scale) : 1.f, rr, ve[e]);
                }
                const int cpe = (col + e) / a.Cw, cce = (col + e) - cpe * a.Cw;
                a.C[(int64_t)cpe * a.sCp + n * a.ldc + cce] = relu ? fmaxf(ve[e], 0.f) : ve[e];
              }
            }
          }
        }
        __syncwarp();
        if (ep) te_st += clock64() - te0;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
      if (ep) atomicAdd(&g_mix_prof[6], (unsigned long long)te_ld), atomicAdd(&g_mix_prof[7], (unsigned long long)te_st);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (prof && t == 0 && blockIdx.x == 0) atomicAdd(&g_mix_prof[9], (unsigned long long)(clock64() - t_start));
  if (warp == WARP_MMA) tmem_dealloc(tmem_base, (uint32_t)(2 * P.tmem_cols));
}

static int tma_stages_for(int BN) {
  const size_t stage = 2 * (size_t)A_TILE + 2 * (size_t)BN * 128;
  const size_t avail = 226 * 1024 - 4 * (size_t)EPI_BYTES - 1024 - 512;
  int s = (int)(avail / stage);
  return std::max(1, std::min(s, 6));
}
static size_t tma_smem_bytes_for(int BN) {
  return (size_t)tma_stages_for(BN) * (2 * A_TILE + 2 * BN * 128) + 4 * (size_t)EPI_BYTES + 1024 + 512;
}

static int stages_for(int BN) {
  const size_t stage = 2 * (size_t)A_TILE + 2 * (size_t)BN * 128;
  int s = (int)((220 * 1024) / stage);
  return std::max(2, std::min(s, 6));
}
static size_t smem_bytes_for(int BN) { return (size_t)stages_for(BN) * (2 * A_TILE + 2 * BN * 128) + 1024 + 256; }

// Box = [64 reduction elements x 128 rows] of fp32, no swizzle, zero fill out of bounds.
// rank 2 when the batch is contiguous (row n = flat index), rank 3 (k, v, b) when every 128-row tile
// stays inside one sample.  Returns false when neither applies (caller uses the register-path kernel).
static bool encode_a_maps(const MixArgs& a, TmaArgs* Q) {
  if (a.N >= ((int64_t)1 << 31)) return false;
  const int64_t V = a.rows_per_batch;
  const int64_t B = (a.N + V - 1) / V;
  bool flat = true;
  for (int p = 0; p < a.P; ++p) flat = flat && (a.a_sB[p] == V * a.a_sV[p] || B == 1);
  if (!flat && (V % BM != 0)) return false;
  Q->rank = flat ? 2 : 3;
  for (int p = 0; p < a.P; ++p) {
    if (a.a_sV[p] < a.Ka) return false;
    uint64_t dims[3] = {1, 1, 1}, strides[2] = {0, 0};
    const uint32_t box[3] = {(uint32_t)BKB, (uint32_t)BM, 1};
    if (flat) {
      dims[0] = (uint64_t)a.Ka, dims[1] = (uint64_t)a.N;
      strides[0] = (uint64_t)a.a_sV[p] * 4;
    } else {
      dims[0] = (uint64_t)a.Ka, dims[1] = (uint64_t)V, dims[2] = (uint64_t)B;
      strides[0] = (uint64_t)a.a_sV[p] * 4, strides[1] = (uint64_t)a.a_sB[p] * 4;
    }
    if (!encode_f32_map(&Q->amap[p], a.A[p], flat ? 2 : 3, dims, strides, box)) return false;
  }
  return true;
}

}  // namespace tc

// Widest column tile (DSW_OPT_MIX_BN: 0 = 256; tuning switch, multiple of 16)
static int mix_bn_max() {
  const int64_t v = g_options[DSW_OPT_MIX_BN].load(std::memory_order_relaxed);
  return (v >= 16 && v <= 256 && v % 16 == 0) ? (int)v : 256;
}

size_t mix_tc_workspace_bytes(int32_t P, int32_t Ka, int32_t Nc) {
  const int BN = std::min(mix_bn_max(), (Nc + 15) / 16 * 16);
  const int n_tiles = (Nc + BN - 1) / BN;
  const int nkb = (Ka + tc::BKB - 1) / tc::BKB;
  return (size_t)n_tiles * P * nkb * 2 * BN * 128 + 256;
}

// `prep` = workspace of mix_tc_workspace_bytes(); nullptr -> unsupported (caller falls back).
// do_prep = false reuses the weight images written by an earlier call with the same B operand
// (the per-chunk calls of one convolution share them).
int launch_mix_tc_ws(const MixArgs& a, void* prep, size_t prep_bytes, bool do_prep, cudaStream_t st) {
  if (!prep || a.Nc < 1 || a.Ka < 1 || a.N < 1) return DSW_ERR_UNSUPPORTED;
  if (prep_bytes < mix_tc_workspace_bytes(a.P, a.Ka, a.Nc)) return DSW_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(prep) & 15) return DSW_ERR_UNSUPPORTED;
  tc::TcArgs P;
  P.m = a;
  P.BN = std::min(mix_bn_max(), (a.Nc + 15) / 16 * 16);
  P.nkb = (a.Ka + tc::BKB - 1) / tc::BKB;
  P.dbg = (int)g_options[DSW_OPT_DEBUG].load(std::memory_order_relaxed) & 0xDF0;
  P.cmode = split_mode();
  int cols = 32;
  while (cols < P.BN) cols <<= 1;
  P.tmem_cols = cols;
  P.bprep = static_cast<const uint8_t*>(prep);
  const int n_tiles = (a.Nc + P.BN - 1) / P.BN;

  if (do_prep) {
    const int64_t chunks = (int64_t)n_tiles * a.P * P.nkb * P.BN * 8;
    DSW_CUDA_TRY(launch_pdl(tc::mix_tc_prep_kernel, dim3((unsigned)ceil_div64(chunks, 256)), dim3(256), 0, st, pdl_enabled(), P,
                            static_cast<uint8_t*>(prep), (int32_t)n_tiles));
    DSW_TRY(check_launch());
  }

  bool vec4 = (a.Ka % 4 == 0);
  for (int p = 0; p < a.P && vec4; ++p)
    vec4 = ((reinterpret_cast<uintptr_t>(a.A[p]) & 15) == 0) && (a.a_sB[p] % 4 == 0) && (a.a_sV[p] % 4 == 0);
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    DSW_CUDA_TRY(cudaGetDevice(&dev));
    DSW_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  P.n_ctiles = n_tiles;
  const int64_t total_tiles = ceil_div64(a.N, tc::BM) * n_tiles;
  dim3 grid((unsigned)std::min<int64_t>(total_tiles, n_sm));

  // ---- TMA-fed kernel: needs a tensor map per A plane ----
  if (vec4 && g_options[DSW_OPT_NO_TMA].load(std::memory_order_relaxed) == 0 && tc::tma_stages_for(P.BN) >= 2) {
    tc::TmaArgs Q;
    Q.tc = P;
    Q.tc.stages = tc::tma_stages_for(P.BN);
    if (tc::encode_a_maps(a, &Q)) {
      const size_t smem = tc::tma_smem_bytes_for(P.BN);
      static PerDeviceOnce attr_tma[2];
      if (a.R != nullptr) {
        DSW_CUDA_TRY(attr_tma[1].max_dynamic_smem(tc::mix_tma_kernel<true>, 227 * 1024));
        DSW_CUDA_TRY(launch_pdl(tc::mix_tma_kernel<true>, grid, dim3(tc::THREADS2), smem, st, pdl_enabled(), Q));
      } else {
        DSW_CUDA_TRY(attr_tma[0].max_dynamic_smem(tc::mix_tma_kernel<false>, 227 * 1024));
        DSW_CUDA_TRY(launch_pdl(tc::mix_tma_kernel<false>, grid, dim3(tc::THREADS2), smem, st, pdl_enabled(), Q));
      }
      return check_launch();
    }
  }

  const size_t smem = tc::smem_bytes_for(P.BN);
  P.stages = tc::stages_for(P.BN);
  static PerDeviceOnce attr_set[2];
  if (vec4) {
    DSW_CUDA_TRY(attr_set[1].max_dynamic_smem(tc::mix_tc_kernel<true>, 227 * 1024));
    tc::mix_tc_kernel<true><<<grid, tc::THREADS2, smem, st>>>(P);
  } else {
    DSW_CUDA_TRY(attr_set[0].max_dynamic_smem(tc::mix_tc_kernel<false>, 227 * 1024));
    tc::mix_tc_kernel<false><<<grid, tc::THREADS2, smem, st>>>(P);
  }
  return check_launch();
}

}  // namespace dsw

extern "C" int dsw_debug_mix_counters(uint64_t* out16, int reset) {
  if (!out16) return DSW_ERR_BAD_ARGUMENT;
  unsigned long long h[16];
  DSW_CUDA_TRY(cudaMemcpyFromSymbol(h, dsw::tc::g_mix_prof, sizeof(h)));
  for (int i = 0; i < 16; ++i) out16[i] = h[i];
  if (reset) {
    unsigned long long z[16] = {};
    DSW_CUDA_TRY(cudaMemcpyToSymbol(dsw::tc::g_mix_prof, z, sizeof(z)));
  }
  return DSW_OK;
}
