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
// CTA = one 128-row x BN-column output tile (BN <= 256, TMEM accumulator 128 lanes x BN columns).
//   warps 0-7  converters: fp32 A rows (global, coalesced float4) -> bf16 hi/lo -> shared memory in the
//              UMMA canonical K-major SWIZZLE_128B layout; afterwards the epilogue
//              (tcgen05.ld -> +bias / ReLU -> global).
//   warp 8     one elected lane: bulk-async copy (TMA, cp.async.bulk) of the pre-split B block into
//              shared memory, tcgen05.mma issue, tcgen05.commit onto the stage's "empty" mbarrier.
// Two-stage ring per CTA; with BN <= 64 two CTAs share an SM so one CTA's epilogue overlaps the other's
// main loop.  B (the weights) is split and laid out as ready-to-copy shared-memory images by a small
// prep kernel once per call.
#include <cuda_bf16.h>

#include <algorithm>

#include "dsw_internal.cuh"

namespace dsw {
namespace tc {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BKB = 64;          // reduction elements per stage: 64 bf16 = one 128-byte swizzle row
constexpr int A_TILE = BM * 128; // bytes of one bf16 A image (hi or lo)
constexpr int CONV_THREADS = 256;
constexpr int THREADS = CONV_THREADS + 32;
constexpr int STAGES = 2;

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
  int32_t tmem_cols;     // power of two >= max(32, BN)
};

// ---------------------------------------------------------------------------------------------
// B preparation: split weights into bf16 hi/lo and write shared-memory images
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mix_tc_prep_kernel(TcArgs P, uint8_t* __restrict__ out, int32_t n_tiles) {
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
// main kernel
// ---------------------------------------------------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(THREADS, 1) mix_tc_kernel(const __grid_constant__ TcArgs P) {
  extern __shared__ uint8_t smem_raw[];
  const MixArgs& a = P.m;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int BN = P.BN;
  const uint32_t b_img = (uint32_t)BN * 128u;
  const uint32_t stage_bytes = 2u * A_TILE + 2u * b_img;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + STAGES * stage_bytes;
  // barrier slots (8 bytes each): a_full[2], b_full[2], empty[2]; then the TMEM address slot
  auto a_full = [&](int s) { return bars + 8u * s; };
  auto b_full = [&](int s) { return bars + 16u + 8u * s; };
  auto empty = [&](int s) { return bars + 32u + 8u * s; };
  const uint32_t tmem_slot = bars + 48u;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * stage_bytes + 48);

  if (t == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(a_full(s), CONV_THREADS);
      mbar_init(b_full(s), 1);
      mbar_init(empty(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int64_t n0 = (int64_t)blockIdx.x * BM;
  const int tile = blockIdx.y;
  const int total_kb = a.P * P.nkb;
  const int last_ksteps = (a.Ka - (P.nkb - 1) * BKB + 15) / 16;  // K-steps (of 16) in a plane's last block

  if (warp < 8) {
    // ================= converters =================
    // thread -> rows (t>>4) + 16*i, i = 0..7; float4 column q = t & 15 (reduction elements 4q..4q+3)
    const int q = t & 15;
    int32_t rb[8], rv[8];
    bool rok[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t n = n0 + (t >> 4) + 16 * i;
      rok[i] = n < a.N;
      const int64_t bb = rok[i] ? n / a.rows_per_batch : 0;
      rb[i] = (int32_t)bb;
      rv[i] = rok[i] ? (int32_t)(n - bb * a.rows_per_batch) : 0;
    }
    float4 cur[8];
    auto load_block = [&](int kbi, float4 (&dst)[8]) {
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
            if (kk < a.Ka) v = __ldg(reinterpret_cast<const float4*>(src));  // Ka % 4 == 0
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
    load_block(0, cur);
    for (int kbi = 0; kbi < total_kb; ++kbi) {
      const int s = kbi & 1;
      if (kbi >= STAGES) {
        mbar_wait(empty(s), ((kbi >> 1) - 1) & 1);
        tc_fence_after();
      }
      uint8_t* Ahi = smem_gen + s * stage_bytes;
      uint8_t* Alo = Ahi + A_TILE;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t row = (t >> 4) + 16 * i;
        __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
        split_bf16(cur[i].x, h0, l0);
        split_bf16(cur[i].y, h1, l1);
        split_bf16(cur[i].z, h2, l2);
        split_bf16(cur[i].w, h3, l3);
        const uint32_t off = swz(row, q >> 1) + (q & 1) * 8;
        *reinterpret_cast<uint2*>(Ahi + off) = make_uint2(pack2(h0, h1), pack2(h2, h3));
        *reinterpret_cast<uint2*>(Alo + off) = make_uint2(pack2(l0, l1), pack2(l2, l3));
      }
      if (kbi + 1 < total_kb) load_block(kbi + 1, cur);  // in flight while the MMAs of this block run
      fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(a_full(s));
    }
    // ================= epilogue =================
    {
      const int last = total_kb - 1;
      mbar_wait(empty(last & 1), (last >> 1) & 1);
      tc_fence_after();
    }
    const int quarter = warp & 3;                   // TMEM lanes 32*quarter .. +31
    const int half = warp >> 2;                     // column half handled by this warp
    const int64_t n = n0 + quarter * 32 + lane;     // output row of this thread
    const int chunks = BN / 16;
    const int c_begin = (chunks * half) / 2, c_end = (chunks * (half + 1)) / 2;
    const bool vec_store = (a.Cw % 4 == 0) && (a.ldc % 4 == 0) && (a.sCp % 4 == 0) &&
                           ((reinterpret_cast<uintptr_t>(a.C) & 15) == 0);
    for (int ch = c_begin; ch < c_end; ++ch) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ch * 16), r);
      tmem_ld_wait();
      if (n >= a.N) continue;
      const int cg0 = tile * BN + ch * 16;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const int cg = cg0 + j;
        if (cg >= a.Nc) break;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[e] = __uint_as_float(r[j + e]);
          if (a.bias && cg + e < a.Nc) v[e] += __ldg(a.bias + cg + e);
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
  } else if (lane == 0) {
    // ================= TMA + MMA issuer (one thread) =================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint8_t* bsrc = P.bprep + (int64_t)tile * total_kb * 2 * b_img;
    for (int kbi = 0; kbi < total_kb; ++kbi) {
      const int s = kbi & 1;
      if (kbi >= STAGES) mbar_wait(empty(s), ((kbi >> 1) - 1) & 1);
      const uint32_t st_base = smem_base + s * stage_bytes;
      mbar_expect_tx(b_full(s), 2u * b_img);
      bulk_copy_g2s(st_base + 2u * A_TILE, bsrc + (int64_t)kbi * 2 * b_img, 2u * b_img, b_full(s));
      mbar_wait(a_full(s), (kbi >> 1) & 1);
      mbar_wait(b_full(s), (kbi >> 1) & 1);
      tc_fence_after();
      const int kb = kbi % P.nkb;
      const int ksteps = (kb == P.nkb - 1) ? last_ksteps : BKB / 16;
      const uint64_t dAh = make_desc(st_base), dAl = make_desc(st_base + A_TILE);
      const uint64_t dBh = make_desc(st_base + 2u * A_TILE), dBl = make_desc(st_base + 2u * A_TILE + b_img);
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 2);  // 32 bytes per K-step, in 16-byte units
        umma_bf16(tmem_base, dAh + adv, dBh + adv, idesc, (kbi | ks) != 0);
        umma_bf16(tmem_base, dAh + adv, dBl + adv, idesc, 1u);
        umma_bf16(tmem_base, dAl + adv, dBh + adv, idesc, 1u);
      }
      umma_commit(empty(s));  // arrives when every MMA issued so far has completed
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

static size_t smem_bytes_for(int BN) { return (size_t)STAGES * (2 * A_TILE + 2 * BN * 128) + 1024 + 64; }

}  // namespace tc

size_t mix_tc_workspace_bytes(int32_t P, int32_t Ka, int32_t Nc) {
  const int BN = std::min(256, (Nc + 15) / 16 * 16);
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
  P.BN = std::min(256, (a.Nc + 15) / 16 * 16);
  P.nkb = (a.Ka + tc::BKB - 1) / tc::BKB;
  int cols = 32;
  while (cols < P.BN) cols <<= 1;
  P.tmem_cols = cols;
  P.bprep = static_cast<const uint8_t*>(prep);
  const int n_tiles = (a.Nc + P.BN - 1) / P.BN;

  if (do_prep) {
    const int64_t chunks = (int64_t)n_tiles * a.P * P.nkb * P.BN * 8;
    tc::mix_tc_prep_kernel<<<(unsigned)ceil_div64(chunks, 256), 256, 0, st>>>(P, static_cast<uint8_t*>(prep), n_tiles);
    DSW_TRY(check_launch());
  }

  bool vec4 = (a.Ka % 4 == 0);
  for (int p = 0; p < a.P && vec4; ++p)
    vec4 = ((reinterpret_cast<uintptr_t>(a.A[p]) & 15) == 0) && (a.a_sB[p] % 4 == 0) && (a.a_sV[p] % 4 == 0);
  const size_t smem = tc::smem_bytes_for(P.BN);
  dim3 grid((unsigned)ceil_div64(a.N, tc::BM), n_tiles);
  static std::atomic<bool> attr_set[2] = {{false}, {false}};
  if (vec4) {
    if (!attr_set[1].exchange(true))
      DSW_CUDA_TRY(cudaFuncSetAttribute(tc::mix_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    tc::mix_tc_kernel<true><<<grid, tc::THREADS, smem, st>>>(P);
  } else {
    if (!attr_set[0].exchange(true))
      DSW_CUDA_TRY(cudaFuncSetAttribute(tc::mix_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    tc::mix_tc_kernel<false><<<grid, tc::THREADS, smem, st>>>(P);
  }
  return check_launch();
}

}  // namespace dsw
