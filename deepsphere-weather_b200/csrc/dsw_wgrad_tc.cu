// tcgen05 weight gradient for sm_100a:
//     dW[f][k][o] = sum_n T_k[n][f] * dY[n][o],      dbias[o] = sum_n dY[n][o]
// (autograd of the channel mix, reference modules/layers.py:176-177, and of the bias add :375).
//
// The reduction runs over the rows n = (b, v), which is the *slow* index of both operands in memory
// ([n][f] and [n][o], channel-last), so both UMMA operands are MN-major: a shared-memory row is one
// n holding 64 consecutive channels (128 bytes of bf16), 8-row groups form the 1024-byte SWIZZLE_128B
// atom, 64-channel blocks are LBO apart.  The global -> shared path therefore keeps the memory order:
// coalesced float4 loads, split into bf16 hi/lo (3-term product, see dsw_mix_tc.cu), 8-byte stores.
//
// CTA = one 128 x BN tile of dW for one row range ("split"): TMEM lanes 0-63 / 64-127 hold two
// (plane k, 64-channel block) half-tiles, BN <= 256 output channels in columns.  16 converter warps
// stream units of 64 rows x 64 channels through a 4-deep register ring (64 KB of loads in flight per
// SM — the small layers are HBM-bound); warp 16 issues the MMAs.  Partials go to the workspace and
// are summed in a fixed order by wgrad_reduce_kernel (deterministic).
#include <cuda_bf16.h>

#include <algorithm>

#include "dsw_internal.cuh"
#include "dsw_tmap.cuh"

namespace dsw {

int launch_wgrad_reduce(const float* partial, int32_t nsplit, int32_t K, int32_t Fin, int32_t Fout, float* dW,
                        float* dbias, cudaStream_t st);

namespace wtc {

constexpr int KB = 64;                 // rows (reduction) per stage
constexpr int BLK = 64 * 128;          // bytes of one 64-row x 64-channel bf16 block image
constexpr int CONV_THREADS = 512;
constexpr int THREADS = CONV_THREADS + 32;
constexpr int STAGES = 2;
constexpr int RING = 4;                // units in flight per thread

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

// ---- CTA pair (cta_group::2) variants: the two CTAs of a cluster share one 256-row MMA ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// arrive on the barrier at the same offset in the pair's leader CTA (rank 0)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAITC_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONEC_%=;\n\t"
      "bra WAITC_%=;\n\t"
      "DONEC_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// MN-major SWIZZLE_128B descriptor: `lbo` = bytes between 64-element MN blocks, `sbo` = bytes between
// 8-row K groups (cute::UMMA::make_umma_desc<Major::MN>: ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO))).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t chunk) {
  return row * 128u + (((chunk ^ (row & 7u)) & 7u) << 4);
}
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

struct WtcArgs {
  WgradArgs w;
  int32_t BN;          // output channels per CTA (multiple of 16, <= 256)
  int32_t nb;          // ceil(BN / 64)
  int32_t ftiles;      // ceil(Fin / 64)
  int32_t tmem_cols;
  int32_t kb_per_split;
  int32_t otiles;      // column tiles per Fout-side plane
  int32_t dbg;         // timing experiments only (DSW_OPT_DEBUG bits 32..256; results become wrong)
  int32_t cmode;       // split_pair mode
  int32_t pp;          // Fout-side planes per CTA tile (TMA kernel only): > 1 = the tile spans pp whole planes of Fout <= 128 columns
};

__global__ void __launch_bounds__(THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WtcArgs P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float bias_acc[256];
  const WgradArgs& a = P.w;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int BN = P.BN, nb = P.nb;
  const int cmode = P.cmode;
  const uint32_t a_img = 2u * BLK;            // one A image (hi or lo): two 64-channel blocks
  const uint32_t b_img = (uint32_t)nb * BLK;  // one B image
  const uint32_t stage_bytes = 2u * a_img + 2u * b_img;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + STAGES * stage_bytes;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 16u + 8u * s; };
  const uint32_t tmem_slot = bars + 32u;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * stage_bytes + 32);

  if (t == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full(s), CONV_THREADS);
      mbar_init(empty(s), 1);
    }
    fence_mbar_init();
  }
  if (t < 256) bias_acc[t] = 0.f;
  if (warp == 16) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int mtile = blockIdx.x, split = blockIdx.z;
  const int kb_plane = blockIdx.y / P.otiles, ntile = blockIdx.y - kb_plane * P.otiles;  // Fout-side plane, column tile
  const float* __restrict__ Yp = a.Y[kb_plane];
  const int n_half = a.Ka * P.ftiles;
  const int64_t r_begin = (int64_t)split * P.kb_per_split * KB;
  const int64_t r_end = (r_begin + (int64_t)P.kb_per_split * KB < a.N) ? r_begin + (int64_t)P.kb_per_split * KB : a.N;
  const int nkb = (r_end > r_begin) ? (int)((r_end - r_begin + KB - 1) / KB) : 0;
  const int U = 2 + nb;  // units per row block: A half 0, A half 1, B blocks
  const int o_base = ntile * BN;

  if (warp < 16) {
    // ================= converters =================
    const int q = t & 15;        // float4 index inside a 64-channel row
    const int r0 = t >> 4;       // rows r0 and r0 + 32 of the unit
    // half-tile sources
    const float* hsrc[2];
    int64_t h_sB[2], h_sV[2];
    int h_valid_f[2];
    bool h_contig[2];
    for (int h = 0; h < 2; ++h) {
      h_contig[h] = true;
      const int ht = mtile * 2 + h;
      if (ht < n_half) {
        const int k = ht / P.ftiles, f0 = (ht - k * P.ftiles) * 64;
        hsrc[h] = a.T[k] + f0;
        h_sB[h] = a.t_sB[k], h_sV[h] = a.t_sV[k];
        h_valid_f[h] = a.Fin - f0;  // channels available from f0
        h_contig[h] = (h_sB[h] == (int64_t)a.rows_per_batch * h_sV[h]);
      } else {
        hsrc[h] = nullptr, h_sB[h] = 0, h_sV[h] = 0, h_valid_f[h] = 0;
      }
    }
    const bool do_bias = (mtile == 0) && (kb_plane == 0) && (a.dbias != nullptr);
    float4 bsum[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bsum[j] = make_float4(0.f, 0.f, 0.f, 0.f);

    const int total_units = nkb * U;
    float4 ring[RING][2];

    auto load_unit = [&](int j, float4 (&dst)[2]) {
      const int kb = j / U, u = j - kb * U;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int64_t n = r_begin + (int64_t)kb * KB + r0 + 32 * i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < r_end) {
          if (u < 2) {
            const int avail = h_valid_f[u] - q * 4;
            if (avail > 0) {
              const float* src;
              if (h_contig[u]) {
                src = hsrc[u] + n * h_sV[u] + q * 4;
              } else {
                const uint32_t bb = (uint32_t)n / (uint32_t)a.rows_per_batch;
                src = hsrc[u] + (int64_t)bb * h_sB[u] + (int64_t)((uint32_t)n - bb * (uint32_t)a.rows_per_batch) * h_sV[u] + q * 4;
              }
              if (avail >= 4 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
                v = __ldg(reinterpret_cast<const float4*>(src));
              } else {
                v.x = __ldg(src);
                if (avail > 1) v.y = __ldg(src + 1);
                if (avail > 2) v.z = __ldg(src + 2);
                if (avail > 3) v.w = __ldg(src + 3);
              }
            }
          } else {
            const int o = o_base + (u - 2) * 64 + q * 4;
            const int avail = a.Fout - o;
            if (avail > 0) {
              const float* src = Yp + n * a.Fout + o;
              if (avail >= 4 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
                v = __ldg(reinterpret_cast<const float4*>(src));
              } else {
                v.x = __ldg(src);
                if (avail > 1) v.y = __ldg(src + 1);
                if (avail > 2) v.z = __ldg(src + 2);
                if (avail > 3) v.w = __ldg(src + 3);
              }
            }
          }
        }
        dst[i] = v;
      }
    };

    auto store_unit = [&](int j, const float4 (&src)[2]) {
      const int kb = j / U, u = j - kb * U;
      const int s = kb & 1;
      if (u == 0 && kb >= STAGES) {
        mbar_wait(empty(s), ((kb >> 1) - 1) & 1);
        tc_fence_after();
      }
      uint8_t* st = smem_gen + s * stage_bytes;
      uint8_t* hi_img = (u < 2) ? st + u * BLK : st + 2u * a_img + (u - 2) * BLK;
      uint8_t* lo_img = (u < 2) ? hi_img + a_img : hi_img + b_img;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint32_t row = r0 + 32 * i;
        uint2 qh, ql;
        split_quad(src[i], qh, ql, cmode);
        const uint32_t off = swz(row, q >> 1) + (q & 1) * 8;
        *reinterpret_cast<uint2*>(hi_img + off) = qh;
        *reinterpret_cast<uint2*>(lo_img + off) = ql;
      }
      if (do_bias && u >= 2) {
        float4& b = bsum[u - 2];
        b.x += src[0].x + src[1].x, b.y += src[0].y + src[1].y;
        b.z += src[0].z + src[1].z, b.w += src[0].w + src[1].w;
      }
      if (u == U - 1) {
        fence_proxy_async();
        mbar_arrive(full(s));
      }
    };

#pragma unroll
    for (int d = 0; d < RING; ++d)
      if (d < total_units) load_unit(d, ring[d]);
    for (int j0 = 0; j0 < total_units; j0 += RING) {
#pragma unroll
      for (int d = 0; d < RING; ++d) {
        const int j = j0 + d;
        if (j < total_units) {
          store_unit(j, ring[d]);
          if (j + RING < total_units) load_unit(j + RING, ring[d]);
        }
      }
    }

    // ---- dbias partial (fp32, CTA-local reduction through shared memory) ----
    if (do_bias) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < nb) {
          atomicAdd(&bias_acc[j * 64 + q * 4 + 0], bsum[j].x);
          atomicAdd(&bias_acc[j * 64 + q * 4 + 1], bsum[j].y);
          atomicAdd(&bias_acc[j * 64 + q * 4 + 2], bsum[j].z);
          atomicAdd(&bias_acc[j * 64 + q * 4 + 3], bsum[j].w);
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(CONV_THREADS));

    // ================= epilogue: TMEM -> partial[split] =================
    const int64_t prow = (int64_t)a.K * a.Fin + 1;
    float* __restrict__ Pp = a.partial + (int64_t)split * prow * a.Fout;
    if (do_bias && t < BN && o_base + t < a.Fout) Pp[(prow - 1) * a.Fout + o_base + t] = bias_acc[t];

    if (nkb > 0) {
      const int last = nkb - 1;
      mbar_wait(empty(last & 1), (last >> 1) & 1);
      tc_fence_after();
    }
    const int quarter = warp & 3, part = warp >> 2;  // 4 lane quarters x 4 column parts
    const int L = quarter * 32 + lane;
    const int h = L >> 6, fl = L & 63;
    const int ht = mtile * 2 + h;
    int64_t m = -1;
    if (ht < n_half) {
      const int k = ht / P.ftiles, f0 = (ht - k * P.ftiles) * 64;
      if (f0 + fl < a.Fin) m = (int64_t)(f0 + fl) * a.K + k + kb_plane;
    }
    const int chunks = BN / 16;
    for (int ch = (chunks * part) / 4; ch < (chunks * (part + 1)) / 4; ++ch) {
      uint32_t r[16];
      if (nkb > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ch * 16), r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = 0u;
      }
      if (m < 0) continue;
      const int o0 = o_base + ch * 16;
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (o0 + e < a.Fout) Pp[m * a.Fout + o0 + e] = __uint_as_float(r[e]);
    }
  } else if (lane == 0) {
    // ================= MMA issuer =================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t lbo = (uint32_t)BLK;  // between 64-channel blocks
    const uint32_t sbo = 1024u;          // between 8-row groups
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb & 1;
      mbar_wait(full(s), (kb >> 1) & 1);
      tc_fence_after();
      const uint32_t st = smem_base + s * stage_bytes;
      const uint32_t Ah = st, Al = st + a_img, Bh = st + 2u * a_img, Bl = Bh + b_img;
      for (int ks = 0; ks < KB / 16; ++ks) {
        const uint32_t adv = (uint32_t)ks * 2048u;  // 16 rows = two 8-row groups
        const uint64_t dAh = make_desc_mn(Ah + adv, lbo, sbo), dAl = make_desc_mn(Al + adv, lbo, sbo);
        const uint64_t dBh = make_desc_mn(Bh + adv, lbo, sbo), dBl = make_desc_mn(Bl + adv, lbo, sbo);
        umma_bf16(tmem_base, dAh, dBh, idesc, (kb | ks) != 0);
        umma_bf16(tmem_base, dAh, dBl, idesc, 1u);
        umma_bf16(tmem_base, dAl, dBh, idesc, 1u);
      }
      umma_commit(empty(s));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// TMA-fed variant (production path for 16-byte aligned operands).
//
// Every unit (64 rows x 64 channels of fp32 = 16 KB) is landed in the stage by one tensor-map TMA box
// issued by a producer lane (zero fill past the row range / channel count), up to S stages ahead of
// the arithmetic; the converters read the raw unit from shared memory and overwrite it in place with
// its bf16 hi (first 8 KB) and lo (second 8 KB) MN-major SWIZZLE_128B images, so consecutive 64-channel
// blocks of one operand are 16 KB apart (the descriptors' LBO).
//
//   warps 0-7 converters (+ epilogue) | warp 8 producer | warp 9 MMA issuer
// ---------------------------------------------------------------------------------------------
constexpr int UNIT = 64 * 64 * 4;       // bytes of one raw unit = its two bf16 images
constexpr int T_CONV = 256;
constexpr int T_THREADS = T_CONV + 64;

struct WtmaArgs {
  WtcArgs c;
  int32_t stages;
  int32_t rank;                       // 2: Fin-side rows are the flat (b, v) index; 3: (f, v, b) coordinates
  CUtensorMap maps[DSW_MAX_K + 1];    // Fin-side planes first (Ka of them), then the Fout-side planes (Kb)
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// PAIR: two CTAs of a cluster (consecutive M tiles, same N tile and row range) issue one
// tcgen05.mma.cta_group::2 of 256 x BN: each CTA stages, converts and feeds its own 128 rows of the M side
// and only HALF of the N-side units — a third less shared-memory traffic per MMA cycle, which is what
// bounds this kernel.  Rank 0 issues the MMAs once both CTAs' converters have arrived on its barrier; the
// commits are multicast to both CTAs.
// Role timing for tuning (DSW_OPT_DEBUG bit 512): cycles summed over CTAs, read by dsw_debug_dense_counters.
__device__ unsigned long long g_wgrad_prof[8];

template <bool PAIR>
__global__ void __launch_bounds__(T_THREADS, 1) wgrad_tma_kernel(const __grid_constant__ WtmaArgs Q) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  __shared__ float bias_acc[256];
  const WtcArgs& P = Q.c;
  const WgradArgs& a = P.w;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int BN = P.BN, nb = P.nb, S = Q.stages;
  const int cmode = P.cmode;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int nbl = PAIR ? nb / 2 : nb;          // N-side units staged by this CTA
  const int ub0 = PAIR ? (int)rank * nbl : 0;  // their first unit index inside the N tile
  const int U = 2 + nbl;  // units per row block: A half 0, A half 1, local B blocks
  const uint32_t stage_bytes = (uint32_t)U * UNIT;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + (uint32_t)S * stage_bytes;
  auto raw_full = [&](int s) { return bars + 8u * s; };
  auto full = [&](int s) { return bars + 8u * (S + s); };
  auto empty = [&](int s) { return bars + 8u * (2 * S + s); };
  const uint32_t tmem_slot = bars + 8u * (3 * S);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (size_t)S * stage_bytes + 8 * (3 * S));

  const int mtile = blockIdx.x, split = blockIdx.z;
  // N tile: either one column tile of one Fout-side plane, or (pp > 1) pp whole planes side by side
  const bool span = P.pp > 1;
  const int Fu = (a.Fout + 63) / 64 * 64;  // plane width in whole units (span: each plane starts a new unit)
  const int kb_plane = span ? blockIdx.y * P.pp : blockIdx.y / P.otiles;
  const int ntile = span ? 0 : blockIdx.y - kb_plane * P.otiles;
  const int n_half = a.Ka * P.ftiles;
  const int64_t r_begin = (int64_t)split * P.kb_per_split * KB;
  const int64_t r_end = (r_begin + (int64_t)P.kb_per_split * KB < a.N) ? r_begin + (int64_t)P.kb_per_split * KB : a.N;
  const int nkb = (r_end > r_begin) ? (int)((r_end - r_begin + KB - 1) / KB) : 0;
  const int o_base = ntile * BN;
  const bool half_ok[2] = {mtile * 2 < n_half, mtile * 2 + 1 < n_half};

  if (t == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(raw_full(s), 1);
      mbar_init(full(s), (PAIR ? 2 : 1) * (T_CONV / 32));  // one elected arrival per converter warp (of both CTAs)
      mbar_init(empty(s), 1);
    }
    fence_mbar_init();
  }
  if (t < 256) bias_acc[t] = 0.f;
  // a missing second half-tile is never loaded: keep its images zero in every stage
  if (!half_ok[1])
    for (int s = 0; s < S; ++s)
      for (int i = t; i < UNIT / 16; i += T_THREADS)
        *reinterpret_cast<uint4*>(smem_gen + (size_t)s * stage_bytes + UNIT + i * 16) = make_uint4(0, 0, 0, 0);
  if (warp == 9) {
    if (PAIR) tmem_alloc2(tmem_slot, (uint32_t)P.tmem_cols);
    else tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anyone arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();  // barriers and TMEM are set up; the operands may still be being written by the previous kernel

  if (warp < 8) {
    // ================= converters =================
    const int q = t & 15, r0 = t >> 4;  // float4 column, rows r0 + 16 i
    const bool do_bias = ((PAIR ? mtile >> 1 : mtile) == 0) && (kb_plane == 0) && (a.dbias != nullptr);
    float4 bsum[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bsum[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % S;
      const uint32_t ph = (uint32_t)(kb / S) & 1;
      uint8_t* st = smem_gen + (size_t)s * stage_bytes;
      const bool prof = (P.dbg & 512) && t == 0;
      long long tp0 = 0;
      if (prof) tp0 = clock64();
      mbar_wait(raw_full(s), ph);
      if (prof) {
        const long long tp1 = clock64();
        atomicAdd(&g_wgrad_prof[0], (unsigned long long)(tp1 - tp0));  // converters waiting for the TMA
        tp0 = tp1;
      }
      if (P.dbg & 128) {  // timing experiment: no conversion
        if (lane == 0) {
          if (PAIR) mbar_arrive_leader(full(s));
          else mbar_arrive(full(s));
        }
        continue;
      }
      float4 v[6][4];
#pragma unroll
      for (int u = 0; u < 6; ++u)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          v[u][i] = (u < U && (u != 1 || half_ok[1]))
                        ? *reinterpret_cast<const float4*>(st + u * UNIT + (r0 + 16 * i) * 256 + q * 16)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
      asm volatile("bar.sync 1, %0;" ::"n"(T_CONV) : "memory");  // every converter has read its share
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        if (u >= U || (u == 1 && !half_ok[1])) continue;
        uint8_t* hi_img = st + u * UNIT;
        uint8_t* lo_img = hi_img + UNIT / 2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t row = r0 + 16 * i;
          uint2 qh, ql;
          split_quad(v[u][i], qh, ql, cmode);
          const uint32_t off = swz(row, q >> 1) + (q & 1) * 8;
          *reinterpret_cast<uint2*>(hi_img + off) = qh;
          *reinterpret_cast<uint2*>(lo_img + off) = ql;
        }
        if (do_bias && u >= 2 && (ub0 + u - 2) * 64 < Fu) {  // (pp > 1: only the units of plane 0 are dy itself)
          float4& b = bsum[u - 2];
#pragma unroll
          for (int i = 0; i < 4; ++i) b.x += v[u][i].x, b.y += v[u][i].y, b.z += v[u][i].z, b.w += v[u][i].w;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(full(s));
        else mbar_arrive(full(s));
      }
      if (prof) {
        atomicAdd(&g_wgrad_prof[1], (unsigned long long)(clock64() - tp0));  // converting one stage
        atomicAdd(&g_wgrad_prof[2], 1ull);                                   // stages
      }
    }

    asm volatile("bar.sync 1, %0;" ::"n"(T_CONV) : "memory");

    // ================= epilogue: TMEM -> partial[split] =================
    const int64_t prow = (int64_t)a.K * a.Fin + 1;
    float* __restrict__ Pp = a.partial + (int64_t)split * prow * a.Fout;

    if (nkb > 0) {
      const int last = nkb - 1;
      mbar_wait(empty(last % S), (uint32_t)(last / S) & 1);
      tc_fence_after();
    }
    // ---- dbias partial: fixed-order reduction of the 16 row groups through the (now idle) first stage ----
    if (do_bias) {
      float4* red = reinterpret_cast<float4*>(smem_gen);  // [r0 (16)][unit (4)][q (16)] float4 = 16 KB
#pragma unroll
      for (int j = 0; j < 4; ++j) red[(r0 * 4 + j) * 16 + q] = bsum[j];
      asm volatile("bar.sync 1, %0;" ::"n"(T_CONV) : "memory");
      const int unit = t >> 6, within = t & 63;  // column t of the N tile = unit, float4 q = within / 4, lane within % 4
      if (t < BN && unit >= ub0 && unit < ub0 + nbl && unit * 64 < Fu && o_base + t < a.Fout) {
        const float* redf = reinterpret_cast<const float*>(red);
        float sum = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) sum += redf[(((g * 4 + (unit - ub0)) * 16) + (within >> 2)) * 4 + (within & 3)];
        Pp[(prow - 1) * a.Fout + o_base + t] = sum;
      }
    }
    const int quarter = warp & 3, part = warp >> 2;  // 4 lane quarters x 2 column parts
    const int L = quarter * 32 + lane;
    const int h = L >> 6, fl = L & 63;
    const int ht = mtile * 2 + h;
    int64_t m = -1;
    if (ht < n_half) {
      const int k = ht / P.ftiles, f0 = (ht - k * P.ftiles) * 64;
      if (f0 + fl < a.Fin) m = (int64_t)(f0 + fl) * a.K + k + kb_plane;
    }
    const int chunks = BN / 16;
    for (int ch = (chunks * part) / 2; ch < (chunks * (part + 1)) / 2; ++ch) {
      uint32_t r[16];
      if (nkb > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ch * 16), r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = 0u;
      }
      if (m < 0) continue;
      // pp > 1: planes start on unit (64-column) boundaries, so a 16-column chunk lies inside one plane
      const int cpl = span ? (ch * 16) / Fu : 0;
      const int o0 = span ? (ch * 16) - cpl * Fu : o_base + ch * 16;
      const int64_t mm = m + cpl;
      if (o0 + 15 < a.Fout && (a.Fout & 3) == 0) {
#pragma unroll
        for (int e = 0; e < 16; e += 4)
          *reinterpret_cast<uint4*>(Pp + mm * a.Fout + o0 + e) = make_uint4(r[e], r[e + 1], r[e + 2], r[e + 3]);
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (o0 + e < a.Fout) Pp[mm * a.Fout + o0 + e] = __uint_as_float(r[e]);
      }
    }
  } else if (warp == 8) {
    // ================= producer (one thread): one TMA box per unit =================
    if (lane == 0) {
      int hk[2], hf[2];
      for (int h = 0; h < 2; ++h) {
        const int ht = mtile * 2 + h;
        hk[h] = half_ok[h] ? ht / P.ftiles : 0;
        hf[h] = half_ok[h] ? (ht - hk[h] * P.ftiles) * 64 : 0;
      }
      const int upp = span ? Fu / 64 : 0;  // units per plane when the tile spans planes
            for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        const uint32_t use = (uint32_t)(kb / S);
        long long tq0 = 0;
        if (P.dbg & 512) tq0 = clock64();
        if (use > 0) mbar_wait(empty(s), (use - 1) & 1);
        if (P.dbg & 512) atomicAdd(&g_wgrad_prof[3], (unsigned long long)(clock64() - tq0));  // producer waiting for a free stage
        const uint32_t st = smem_base + s * stage_bytes;
        const int64_t n0 = r_begin + (int64_t)kb * KB;
        const uint32_t tx_now = (P.dbg & 32 ? 0u : (uint32_t)(half_ok[1] ? 2 : 1) * UNIT) + (P.dbg & 64 ? 0u : (uint32_t)nbl * UNIT);
        if (tx_now == 0) {
          mbar_arrive(raw_full(s));
          continue;
        }
        mbar_expect_tx(raw_full(s), tx_now);
        for (int h = 0; h < 2; ++h) {
          if (!half_ok[h] || (P.dbg & 32)) continue;
          if (Q.rank == 2) {
            tma_load_2d(st + h * UNIT, &Q.maps[hk[h]], hf[h], (int)n0, raw_full(s));
          } else {
            const int bb = (int)(n0 / a.rows_per_batch);
            tma_load_3d(st + h * UNIT, &Q.maps[hk[h]], hf[h], (int)(n0 - (int64_t)bb * a.rows_per_batch), bb, raw_full(s));
          }
        }
        for (int j = 0; j < ((P.dbg & 64) ? 0 : nbl); ++j) {
          const int g = ub0 + j;  // unit index inside the N tile
          const int pl = span ? g / upp : 0;
          const int col = span ? (g - pl * upp) * 64 : o_base + 64 * g;
          tma_load_2d(st + (2 + j) * UNIT, &Q.maps[a.Ka + kb_plane + pl], col, (int)n0, raw_full(s));
        }
      }
    }
  } else if (lane == 0 && rank == 0) {
    // ================= MMA issuer (pair: the leader CTA only) =================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
    const uint32_t lbo = (uint32_t)UNIT;  // between 64-channel blocks (each unit = [hi | lo])
    const uint32_t sbo = 1024u;           // between 8-row groups
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % S;
      long long tm0 = 0;
      if (P.dbg & 512) tm0 = clock64();
      if (PAIR) mbar_wait_cluster(full(s), (uint32_t)(kb / S) & 1);
      else mbar_wait(full(s), (uint32_t)(kb / S) & 1);
      if (P.dbg & 512) atomicAdd(&g_wgrad_prof[4], (unsigned long long)(clock64() - tm0));  // MMA issuer waiting for operands
      tc_fence_after();
      const uint32_t st = smem_base + s * stage_bytes;
      const uint32_t Ah = st, Al = st + UNIT / 2, Bh = st + 2u * UNIT, Bl = Bh + UNIT / 2;
      for (int ks = 0; ks < ((P.dbg & 256) ? 0 : KB / 16); ++ks) {
        const uint32_t adv = (uint32_t)ks * 2048u;  // 16 rows = two 8-row groups
        const uint64_t dAh = make_desc_mn(Ah + adv, lbo, sbo), dAl = make_desc_mn(Al + adv, lbo, sbo);
        const uint64_t dBh = make_desc_mn(Bh + adv, lbo, sbo), dBl = make_desc_mn(Bl + adv, lbo, sbo);
        if (PAIR) {
          umma_bf16_pair(tmem_base, dAh, dBh, idesc, (kb | ks) != 0);
          umma_bf16_pair(tmem_base, dAh, dBl, idesc, 1u);
          umma_bf16_pair(tmem_base, dAl, dBh, idesc, 1u);
        } else {
          umma_bf16(tmem_base, dAh, dBh, idesc, (kb | ks) != 0);
          umma_bf16(tmem_base, dAh, dBl, idesc, 1u);
          umma_bf16(tmem_base, dAl, dBh, idesc, 1u);
        }
      }
      if (PAIR) umma_commit_pair(empty(s));
      else umma_commit(empty(s));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // both CTAs are done with the shared accumulator / the peer's operands
  if (warp == 9) {
    if (PAIR) tmem_dealloc2(tmem_base, (uint32_t)P.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  }
}

static int tma_stages_for(int nb) {
  const size_t stage = (size_t)(2 + nb) * UNIT;
  const int s = (int)((224 * 1024) / stage);
  return std::max(1, std::min(s, 4));
}
static size_t tma_smem_bytes_for(int nb) { return (size_t)tma_stages_for(nb) * (2 + nb) * UNIT + 1024 + 256; }

// Tensor maps: box = [64 channels x 64 rows] of fp32.  Returns false when the operands do not allow it.
static bool encode_maps(const WgradArgs& a, WtmaArgs* Q) {
  if ((a.Fin & 3) || (a.Fout & 3)) return false;
  const int64_t V = a.rows_per_batch;
  const int64_t B = (a.N + V - 1) / V;
  bool flat = true;
  for (int k = 0; k < a.Ka; ++k) {
    if ((reinterpret_cast<uintptr_t>(a.T[k]) & 15) || (a.t_sV[k] & 3) || (a.t_sB[k] & 3) || a.t_sV[k] < a.Fin) return false;
    flat = flat && (a.t_sB[k] == V * a.t_sV[k] || B == 1);
  }
  if (!flat && (V % KB != 0)) return false;
  Q->rank = flat ? 2 : 3;
  const uint32_t box[3] = {64u, (uint32_t)KB, 1u};
  for (int k = 0; k < a.Ka; ++k) {
    uint64_t dims[3] = {1, 1, 1}, strides[2] = {0, 0};
    dims[0] = (uint64_t)a.Fin;
    if (flat) {
      dims[1] = (uint64_t)a.N, strides[0] = (uint64_t)a.t_sV[k] * 4;
    } else {
      dims[1] = (uint64_t)V, dims[2] = (uint64_t)B;
      strides[0] = (uint64_t)a.t_sV[k] * 4, strides[1] = (uint64_t)a.t_sB[k] * 4;
    }
    if (!encode_f32_map(&Q->maps[k], a.T[k], flat ? 2 : 3, dims, strides, box)) return false;
  }
  for (int k = 0; k < a.Kb; ++k) {
    if (reinterpret_cast<uintptr_t>(a.Y[k]) & 15) return false;
    const uint64_t dims[2] = {(uint64_t)a.Fout, (uint64_t)a.N};
    const uint64_t strides[1] = {(uint64_t)a.Fout * 4};
    if (!encode_f32_map(&Q->maps[a.Ka + k], a.Y[k], 2, dims, strides, box)) return false;
  }
  return true;
}

static size_t smem_bytes_for(int nb) { return (size_t)STAGES * (2 * 2 * BLK + 2 * nb * BLK) + 1024 + 64; }

}  // namespace wtc

// Planes per N tile.  The adjoint form (Kb = K planes on the Fout side) with narrow planes would give
// N tiles of only 64 / 128 columns — half or a quarter of the MMA work per byte staged; the TMA kernel
// then lets one tile span 256 / ceil64(Fout) whole planes (zero-filled past Fout inside each plane).
static int wgrad_span_planes(int32_t Kb, int32_t Fin, int32_t Fout) {
  if (Kb < 2 || Fout > 128 || (Fout & 3) || (Fin & 3)) return 1;
  if (g_options[DSW_OPT_NO_TMA].load(std::memory_order_relaxed) != 0) return 1;
  int pp = 256 / ((Fout + 63) / 64 * 64);
  while (pp > 1 && Kb % pp) pp >>= 1;
  return pp;
}

static void wgrad_tc_geometry(int64_t N, int32_t Ka, int32_t Kb, int32_t Fin, int32_t Fout, int& BN, int& ntiles,
                              int& mtiles, int& nsplit, int& kb_per_split) {
  const int pp = wgrad_span_planes(Kb, Fin, Fout);
  BN = pp > 1 ? pp * ((Fout + 63) / 64 * 64) : std::min(256, (Fout + 15) / 16 * 16);
  ntiles = pp > 1 ? Kb / pp : Kb * ((Fout + BN - 1) / BN);
  const int ftiles = (Fin + 63) / 64;
  mtiles = (Ka * ftiles + 1) / 2;
  const int64_t total_kb = (N + wtc::KB - 1) / wtc::KB;
  int64_t ns = std::max<int64_t>(1, 148 / ((int64_t)mtiles * ntiles));
  ns = std::min<int64_t>(ns, std::max<int64_t>(1, total_kb / 4));
  kb_per_split = (int)((total_kb + ns - 1) / ns);
  nsplit = (int)((total_kb + kb_per_split - 1) / kb_per_split);
}

int wgrad_tc_nsplit(int64_t N, int32_t Ka, int32_t Kb, int32_t Fin, int32_t Fout) {
  int BN, ntiles, mtiles, nsplit, kbps;
  wgrad_tc_geometry(N, Ka, Kb, Fin, Fout, BN, ntiles, mtiles, nsplit, kbps);
  return nsplit;
}

size_t wgrad_tc_partial_bytes(int64_t N, int32_t Ka, int32_t Kb, int32_t Fin, int32_t Fout) {
  return (size_t)wgrad_tc_nsplit(N, Ka, Kb, Fin, Fout) * ((size_t)Ka * Kb * Fin + 1) * Fout * sizeof(float);
}

// Writes wgrad_tc_nsplit() partials at a.partial; the caller sums them with launch_wgrad_reduce.
int launch_wgrad_tc(const WgradArgs& a, size_t partial_bytes, cudaStream_t st) {
  if (!a.partial) return DSW_ERR_UNSUPPORTED;
  if (partial_bytes < wgrad_tc_partial_bytes(a.N, a.Ka, a.Kb, a.Fin, a.Fout)) return DSW_ERR_UNSUPPORTED;
  if (a.N >= (int64_t)1 << 31) return DSW_ERR_UNSUPPORTED;
  wtc::WtcArgs P;
  P.w = a;
  int ntiles, mtiles, nsplit;
  wgrad_tc_geometry(a.N, a.Ka, a.Kb, a.Fin, a.Fout, P.BN, ntiles, mtiles, nsplit, P.kb_per_split);
  P.pp = wgrad_span_planes(a.Kb, a.Fin, a.Fout);
  P.otiles = P.pp > 1 ? 1 : ntiles / a.Kb;
  P.dbg = (int)g_options[DSW_OPT_DEBUG].load(std::memory_order_relaxed) & 0x3E0;
  P.cmode = split_mode();
  P.w.nsplit = nsplit;
  P.nb = (P.BN + 63) / 64;
  P.ftiles = (a.Fin + 63) / 64;
  int cols = 32;
  while (cols < P.BN) cols <<= 1;
  P.tmem_cols = cols;

  dim3 grid(mtiles, ntiles, nsplit);
  if (g_options[DSW_OPT_NO_TMA].load(std::memory_order_relaxed) == 0 && wtc::tma_stages_for(P.nb) >= 2) {
    wtc::WtmaArgs Q;
    Q.c = P;
    Q.stages = wtc::tma_stages_for(P.nb);
    if (wtc::encode_maps(a, &Q)) {
      // CTA pairs when consecutive M tiles exist in pairs and the N tile splits into two halves of whole units
      // (measured 5-50 % SLOWER than single CTAs — each SM's shared memory still serves its half of B to both
      // tensor cores, so the per-SM read traffic does not drop, and the cross-CTA barriers add latency — hence
      // opt-in only: DSW_OPT_WGRAD_PAIR = 1)
      const bool pair = (mtiles % 2 == 0) && (P.nb % 2 == 0) && g_options[DSW_OPT_WGRAD_PAIR].load(std::memory_order_relaxed) == 1;
      static PerDeviceOnce attr_tma[2];
      if (pair) {
        const int nbl = P.nb / 2;
        Q.stages = wtc::tma_stages_for(nbl);
        DSW_CUDA_TRY(attr_tma[1].max_dynamic_smem(wtc::wgrad_tma_kernel<true>, 226 * 1024));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid, cfg.blockDim = dim3(wtc::T_THREADS), cfg.dynamicSmemBytes = wtc::tma_smem_bytes_for(nbl), cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr, cfg.numAttrs = 1;
        DSW_CUDA_TRY(cudaLaunchKernelEx(&cfg, wtc::wgrad_tma_kernel<true>, Q));
        return check_launch();
      }
      DSW_CUDA_TRY(attr_tma[0].max_dynamic_smem(wtc::wgrad_tma_kernel<false>, 226 * 1024));
      DSW_CUDA_TRY(launch_pdl(wtc::wgrad_tma_kernel<false>, grid, dim3(wtc::T_THREADS), wtc::tma_smem_bytes_for(P.nb), st,
                              pdl_enabled(), Q));
      return check_launch();
    }
  }
  if (P.pp > 1) return launch_wgrad_simt(P.w, st);  // the register-path kernel does not span planes (same partial layout)
  const size_t smem = wtc::smem_bytes_for(P.nb);
  static PerDeviceOnce attr_set;
  DSW_CUDA_TRY(attr_set.max_dynamic_smem(wtc::wgrad_tc_kernel, (int)wtc::smem_bytes_for(4)));  // + 1 KB static bias_acc <= 227 KB
  wtc::wgrad_tc_kernel<<<grid, wtc::THREADS, smem, st>>>(P);
  return check_launch();
}

}  // namespace dsw

extern "C" int dsw_debug_dense_counters(uint64_t* out8, int reset) {
  if (!out8) return DSW_ERR_BAD_ARGUMENT;
  unsigned long long h[8];
  DSW_CUDA_TRY(cudaMemcpyFromSymbol(h, dsw::wtc::g_wgrad_prof, sizeof(h)));
  for (int i = 0; i < 8; ++i) out8[i] = h[i];
  if (reset) {
    unsigned long long z[8] = {};
    DSW_CUDA_TRY(cudaMemcpyToSymbol(dsw::wtc::g_wgrad_prof, z, sizeof(z)));
  }
  return DSW_OK;
}
