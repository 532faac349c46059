// Fused multi-hop chain kernel: every sparse hop of one Chebyshev recurrence in ONE persistent launch.
//
// The K-1 hops of  x_k = 2 L x_{k-1} - x_{k-2}  (reference modules/layers.py:163-169), of its adjoint and of
// the Clenshaw form are each  O_j = alpha_j * (A . X_j) + beta_j * Z_j + G_j  with X_j = O_{j-1}.  Launched one
// kernel per hop, every plane is written to HBM by hop j and read back (twice) by hops j+1 and j+2: 8 planes of
// traffic for a K = 4 recurrence whose algorithmic traffic is 4.  This kernel walks the whole chain in an order
// that keeps the planes it still needs inside the 126 MB L2:
//
//   work item  = (sample group g, hop j, tile t, sample s of the group, 64-channel slab)
//   claim order = that tuple in lexicographic order, taken from ONE global counter by the teams of a persistent
//                 grid (one CTA per SM, CH_TEAMS independent 64-thread teams per CTA)
//   group size  = as many samples as keep three planes of the group (gather source, the k-2 term, the output)
//                 inside the L2 budget, but enough for a hop pass to hold more items than are in flight.
//
// Dependencies instead of grid barriers: item (b, j, t) gathers rows of O_{j-1} that lie in a fixed, small set of
// tiles deps(t) (plan: tdep_ptr / tdep_idx, own tile included).  Every finished item publishes
// flag[b][slab][t] = epoch * 16 + j + 1 (release); an item polls the flags of its deps (acquire) before it issues
// its transfers.  Claims are handed out in order and an item only ever waits for smaller claims, so the grid
// cannot deadlock whatever its size.  The epoch lives in device memory and is bumped by the last CTA to leave,
// which also zeroes the claim counter: no memset launches, and captured CUDA graphs replay correctly.
//
// Per item a team stages, with one mbarrier: the tile's entry-major weight / offset panels (two bulk copies) and
// the distinct source rows the tile gathers (one tensor-map box per run of consecutive rows).  The arithmetic is
// the register-tiled loop of hop_team_kernel (dsw_spmm.cu): 4 lanes own a row-block of 4 rows x 64 channels,
// every 16-byte shared-memory read feeds 4 packed FMAs.
#include <algorithm>

#include "dsw_internal.cuh"
#include "dsw_tmap.cuh"

namespace dsw {

constexpr int CH_TEAMS = 4;
constexpr int CH_TEAM_THREADS = DSW_TILE_BLOCKS * 4;  // 4 lanes per row-block
constexpr int CH_THREADS = CH_TEAMS * CH_TEAM_THREADS;

struct ChainPlan {
  const int32_t* blkptr;
  const int32_t* tp_ptr;
  const float4* tp_val;
  const uint32_t* tp_off;
  const int32_t* tile_ptr;
  const int32_t* tpc_ptr;
  const int32_t* tpc_row;
  const uint32_t* tpc_meta;
  const int32_t* tdep_ptr;
  const int32_t* tdep_idx;
  int32_t n_blocks, n_rows, n_tiles, cap_len, cap_rows;
  int32_t n_hops, n_slabs, F;
  int32_t S, S_last, n_full_groups;  // samples per group, samples of the ragged last group (0 = none)
  int32_t items_full;                // items of a full group = n_hops * n_tiles * S * n_slabs
  int32_t total_items;
  int32_t* sync;                     // this launch's set: [0] claim, [1] CTAs gone, [2] epoch, flags from DSW_CHAIN_HDR
  int32_t n_ctas;
  uint32_t team_stride;              // bytes of shared memory per team
  int32_t debug_skip;                // timing experiments only: 2 = skip the entry loop
};

struct ChainArgs {
  ChainHop h[DSW_CHAIN_MAX_HOPS];
};

struct ChainMaps {
  CUtensorMap m[DSW_CHAIN_MAX_HOPS][8];  // box rows 1, 2, 4, .. 128 over hop j's gather source [B][V][F]
};

__device__ unsigned long long g_chain_prof[8];

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ int32_t ld_acquire(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int32_t* p, int32_t v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void team_bar(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(CH_TEAM_THREADS) : "memory");
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void fma4(float4& acc, float w, const float4& x) {
  const float2 ww = make_float2(w, w);
  float2 lo = __ffma2_rn(ww, make_float2(x.x, x.y), make_float2(acc.x, acc.y));
  float2 hi = __ffma2_rn(ww, make_float2(x.z, x.w), make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void fma_step(float4 (&acc)[4][4], const float4& w, const float4 (&x)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    fma4(acc[0][j], w.x, x[j]);
    fma4(acc[1][j], w.y, x[j]);
    fma4(acc[2][j], w.z, x[j]);
    fma4(acc[3][j], w.w, x[j]);
  }
}

// What a team needs to know about its current item (written by the team's lane 0, read after a team barrier).
struct ItemDesc {
  int32_t idx;   // claim index (>= total_items: nothing left)
  int32_t hop, tile, b, slab;
  int32_t pad[3];
};

}  // namespace

__global__ void __launch_bounds__(CH_THREADS, 1)
    hop_chain_kernel(const ChainPlan P, const __grid_constant__ ChainArgs A, const __grid_constant__ ChainMaps maps) {
  extern __shared__ __align__(1024) uint8_t ch_smem[];
  const int tid = threadIdx.x;
  const int team = tid / CH_TEAM_THREADS;
  const int tt = tid - team * CH_TEAM_THREADS;

  // shared memory: [team 0: rows | weights | offsets] ... [team 3] | mbarriers | item descriptors | epoch
  uint8_t* xs = ch_smem + (size_t)team * P.team_stride;
  const int cap_steps = P.cap_len + DSW_PANEL_PAD;
  const float4* s_val = reinterpret_cast<const float4*>(xs + (size_t)P.cap_rows * 256);
  const uint32_t* s_off = reinterpret_cast<const uint32_t*>(xs + (size_t)P.cap_rows * 256 + (size_t)cap_steps * DSW_TILE_BLOCKS * 16);
  uint8_t* tail = ch_smem + (size_t)CH_TEAMS * P.team_stride;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(tail);
  ItemDesc* s_item = reinterpret_cast<ItemDesc*>(tail + 64);
  int32_t* s_epoch = reinterpret_cast<int32_t*>(tail + 64 + CH_TEAMS * sizeof(ItemDesc));
  const uint32_t xs_u32 = smem_u32(xs);
  const uint32_t bar = smem_u32(s_bar + team);

  pdl_trigger();
  if (tid < CH_TEAMS) mbar_init(smem_u32(s_bar + tid), 1);
  if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  // everything the previous kernel of the stream wrote (operands, and — if the ring of sets has wrapped — this set's
  // sync words) is visible from here on
  pdl_wait();
  if (tid == 0) *s_epoch = *reinterpret_cast<volatile int32_t*>(P.sync + 2);
  __syncthreads();
  const int32_t base = (*s_epoch + 1) * 16;  // flag value of "hop j done" = base + j + 1
  int32_t* const claim = P.sync;
  int32_t* const flags = P.sync + DSW_CHAIN_HDR;

  // ---- staging (the team's first warp) ----
  // Decodes the claim, waits (or just checks) for the tiles the item gathers from, then issues the transfers.
  // Returns false when `blocking` is false and a dependency is still open (nothing issued then).
  auto stage = [&](int32_t idx, bool blocking) -> bool {
    int32_t g, r, Sg;
    if (idx < P.n_full_groups * P.items_full) {
      g = idx / P.items_full, r = idx - g * P.items_full, Sg = P.S;
    } else {
      g = P.n_full_groups, r = idx - g * P.items_full, Sg = P.S_last;
    }
    const int32_t per_tile = Sg * P.n_slabs, per_hop = P.n_tiles * per_tile;
    const int32_t hop = r / per_hop;
    r -= hop * per_hop;
    const int32_t tile = r / per_tile;
    r -= tile * per_tile;
    const int32_t s = r / P.n_slabs, slab = r - s * P.n_slabs;
    const int32_t b = g * P.S + s;
    const ChainHop& H = A.h[hop];
    if (H.dep) {
      const int32_t d0 = __ldg(P.tdep_ptr + tile), nd = __ldg(P.tdep_ptr + tile + 1) - d0;
      const int32_t* f = flags + ((int64_t)b * P.n_slabs + slab) * P.n_tiles;
      const int32_t need = base + hop;  // hop - 1 finished
      const int32_t dep = tt < nd ? __ldg(P.tdep_idx + d0 + tt) : -1;
      bool ok = dep < 0 || ld_acquire(f + dep) - need >= 0;
      if (!blocking) {
        if (!__all_sync(0xffffffffu, ok)) return false;
      } else {
        while (!ok) {
          __nanosleep(100);
          ok = ld_acquire(f + dep) - need >= 0;
        }
        __syncwarp();
      }
    }
    const int32_t t0 = __ldg(P.tp_ptr + tile), steps = __ldg(P.tp_ptr + tile + 1) - t0;
    const int32_t nrows = __ldg(P.tile_ptr + tile + 1) - __ldg(P.tile_ptr + tile);
    const int32_t pc0 = __ldg(P.tpc_ptr + tile), npieces = __ldg(P.tpc_ptr + tile + 1) - pc0;
    // the source rows were written through the generic proxy (by other SMs), the buffer was read through it
    asm volatile("fence.proxy.async;" ::: "memory");
    if (tt == 0) {
      ItemDesc d;
      d.idx = idx, d.hop = hop, d.tile = tile, d.b = b, d.slab = slab;
      s_item[team] = d;
      mbar_expect_tx(bar, (uint32_t)nrows * 256u + (uint32_t)steps * (DSW_TILE_BLOCKS * 20u));
      bulk_g2s(smem_u32(s_val), P.tp_val + (size_t)t0 * DSW_TILE_BLOCKS, (uint32_t)steps * (DSW_TILE_BLOCKS * 16u), bar);
      bulk_g2s(smem_u32(s_off), P.tp_off + (size_t)t0 * DSW_TILE_BLOCKS, (uint32_t)steps * (DSW_TILE_BLOCKS * 4u), bar);
    }
    __syncwarp();
    for (int i = tt; i < npieces; i += 32) {
      const uint32_t meta = __ldg(P.tpc_meta + pc0 + i);
      tma_load_3d(xs_u32 + (meta >> 8) * 256u, &maps.m[hop][meta & 7u], slab * 64, __ldg(P.tpc_row + pc0 + i), b, bar);
    }
    return true;
  };

  const int slot = tt >> 2, lq = tt & 3, par = slot & 1;
  const uint32_t cA = ((uint32_t)(lq * 16) ^ (uint32_t)(par * 64));
  const uint32_t cB = ((uint32_t)(lq * 16 + 64) ^ (uint32_t)(par * 64));
  const uint8_t* xA = xs + cA;
  const uint8_t* xB = xs + cB;
  const uint32_t* po = s_off + slot;
  const float4* pw = s_val + slot;
  int ch[4];
  ch[0] = (int)(cA >> 2), ch[1] = (int)(cB >> 2), ch[2] = ch[0] + 32, ch[3] = ch[1] + 32;

  // first item
  int32_t next = P.total_items;  // lane 0 of the team: the claim after the current one
  if (tt < 32) {
    int32_t first = 0;
    if (tt == 0) first = atomicAdd(claim, 1);
    first = __shfl_sync(0xffffffffu, first, 0);
    if (first < P.total_items) {
      stage(first, true);
    } else if (tt == 0) {
      s_item[team].idx = first;
    }
  }
  uint32_t phase = 0;
  while (true) {
    team_bar(team);  // (A) the descriptor of the current item is visible
    const ItemDesc d = s_item[team];
    if (d.idx >= P.total_items) break;
    if (tt == 0) next = atomicAdd(claim, 1);  // consumed after the entry loop, when the round trip is long over
    const ChainHop& H = A.h[d.hop];
    const int blk = d.tile * DSW_TILE_BLOCKS + slot;
    const bool active = blk < P.n_blocks;
    const int slab_f = min(64, P.F - d.slab * 64);
    int my_len = 0;
    if (active) my_len = __ldg(P.blkptr + blk + 1) - __ldg(P.blkptr + blk);
    const bool prof = (P.debug_skip == 4) && (tt == 0);
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (prof) c0 = clock64();

    // Accumulators start at (beta * Z + G) / alpha (alpha is 1 or 2: exact); the loads land while the tile does.
    float4 acc[4][4];
    {
      const float inv_alpha = 1.f / H.alpha;
      const float zs = H.beta * inv_alpha;
      bool ok[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) ok[r][j] = active && (blk * 4 + r) < P.n_rows && ch[j] < slab_f;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (H.Z != nullptr && ok[r][j])
            acc[r][j] = ldcg4(H.Z + d.b * H.z_sB + (int64_t)(blk * 4 + r) * H.z_sV + d.slab * 64 + ch[j]);
        }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[r][j] = make_float4(acc[r][j].x * zs, acc[r][j].y * zs, acc[r][j].z * zs, acc[r][j].w * zs);
      if (H.G != nullptr) {
        float4 g[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            g[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok[r][j]) g[r][j] = ldcg4(H.G + d.b * H.g_sB + (int64_t)(blk * 4 + r) * H.g_sV + d.slab * 64 + ch[j]);
          }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            acc[r][j] = make_float4(fmaf(g[r][j].x, inv_alpha, acc[r][j].x), fmaf(g[r][j].y, inv_alpha, acc[r][j].y),
                                    fmaf(g[r][j].z, inv_alpha, acc[r][j].z), fmaf(g[r][j].w, inv_alpha, acc[r][j].w));
      }
    }
    const int wlen = (__reduce_max_sync(0xffffffffu, my_len) + 1) & ~1;
    if (prof) c1 = clock64();
    mbar_wait(bar, phase);
    phase ^= 1u;
    if (prof) c2 = clock64();

    {
      auto load_x = [&](uint32_t o, float4(&x)[4]) {
        x[0] = *reinterpret_cast<const float4*>(xA + o);
        x[1] = *reinterpret_cast<const float4*>(xB + o);
        x[2] = *reinterpret_cast<const float4*>(xA + o + 128);
        x[3] = *reinterpret_cast<const float4*>(xB + o + 128);
      };
      // software pipeline: offsets two steps ahead, weights / values one step ahead (the panels end in zero steps)
      float4 x0[4], x1[4], w0, w1;
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
    team_bar(team);  // (B) every lane is done with the staged rows and panels
    if (prof) c3 = clock64();
    // The next item's transfers start now if its dependencies are already met (they almost always are); they then
    // overlap this item's stores.  A blocking wait must not happen before this item's own flag is out (the next
    // item may depend on it).
    bool staged = false;
    int32_t nxt = 0;
    if (tt < 32) {
      nxt = __shfl_sync(0xffffffffu, next, 0);
      if (nxt < P.total_items) staged = stage(nxt, false);
    }
    // ---- epilogue: O = alpha * acc ----
    if (active) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int row = blk * 4 + r;
        if (row >= P.n_rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (ch[j] >= slab_f) continue;
          float4 o = make_float4(H.alpha * acc[r][j].x, H.alpha * acc[r][j].y, H.alpha * acc[r][j].z, H.alpha * acc[r][j].w);
          if (H.act) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
          *reinterpret_cast<float4*>(H.O + d.b * H.o_sB + (int64_t)row * H.o_sV + d.slab * 64 + ch[j]) = o;
        }
      }
    }
    team_bar(team);  // (C) every store of the item has been issued
    if (tt < 32) {
      if (tt == 0) {
        __threadfence();
        st_release(flags + ((int64_t)d.b * P.n_slabs + d.slab) * P.n_tiles + d.tile, base + d.hop + 1);
      }
      __syncwarp();
      if (nxt < P.total_items) {
        if (!staged) stage(nxt, true);
      } else if (tt == 0) {
        s_item[team].idx = nxt;
      }
    }
    if (prof) {
      const long long c4 = clock64();
      atomicAdd(&g_chain_prof[0], (unsigned long long)(c1 - c0));  // Z / G loads
      atomicAdd(&g_chain_prof[1], (unsigned long long)(c2 - c1));  // wait for the transfers
      atomicAdd(&g_chain_prof[2], (unsigned long long)(c3 - c2));  // entry loop
      atomicAdd(&g_chain_prof[3], (unsigned long long)(c4 - c3));  // stores, flag, next staging
      atomicAdd(&g_chain_prof[4], 1ull);
    }
  }

  // Leaving protocol: the last CTA resets the claim counter and opens the next epoch of this set.
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const int32_t gone = atomicAdd(P.sync + 1, 1);
    if (gone == P.n_ctas - 1) {
      P.sync[0] = 0;
      P.sync[1] = 0;
      *reinterpret_cast<volatile int32_t*>(P.sync + 2) = *s_epoch + 1;
      __threadfence();
    }
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int sm_count() {
  static std::atomic<int> cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// Is the fused kernel usable for these hops?  (Otherwise the caller launches them one by one.)
static bool chain_ok(const dsw_csr& A, const dsw_rb& rb, const ChainHop* h, int n, int32_t B, int32_t F) {
  if (g_options[DSW_OPT_NO_CHAIN].load(std::memory_order_relaxed) == 1) return false;
  if (g_options[DSW_OPT_HOP_KERNEL].load(std::memory_order_relaxed) != 0) return false;
  if (g_options[DSW_OPT_NO_TMA].load(std::memory_order_relaxed) != 0) return false;
  if (rb.R != 4 || rb.n_tiles <= 0 || rb.perm || !rb.chain_sync || rb.tile_deps_max <= 0) return false;
  if (A.n_rows != A.n_cols || F % 4 || B <= 0 || B > 65535) return false;
  const int64_t small_f_opt = g_options[DSW_OPT_HOP_SMALL_F].load(std::memory_order_relaxed);
  if (F <= (small_f_opt > 0 ? (int)small_f_opt : 8)) return false;
  if (rb.tile_pieces_max <= 0 || rb.tile_pieces_max > rb.tile_rows_max) return false;
  for (int j = 0; j < n; ++j) {
    const ChainHop& c = h[j];
    if (!c.X || !c.O || !aligned16(c.X) || !aligned16(c.O) || (c.x_sB | c.x_sV | c.o_sB | c.o_sV) % 4 || c.x_sV < F) return false;
    if (c.Z && (!aligned16(c.Z) || (c.z_sB | c.z_sV) % 4)) return false;
    if (c.G && (!aligned16(c.G) || (c.g_sB | c.g_sV) % 4)) return false;
    if (!(c.alpha == 1.f || c.alpha == 2.f)) return false;
    // hop j gathers from hop j-1's output
    if (j > 0 && (c.X != h[j - 1].O || c.x_sB != h[j - 1].o_sB || c.x_sV != h[j - 1].o_sV)) return false;
  }
  return true;
}

static int launch_chain_fused(const dsw_csr& A, const dsw_rb& rb, const ChainHop* hops, int n, int32_t B0, int32_t Bn, int32_t F,
                              cudaStream_t st) {
  ChainPlan P{};
  P.blkptr = rb.blkptr, P.tp_ptr = rb.tp_ptr, P.tp_val = rb.tp_val, P.tp_off = rb.tp_off;
  P.tile_ptr = rb.tile_ptr, P.tpc_ptr = rb.tpc_ptr, P.tpc_row = rb.tpc_row, P.tpc_meta = rb.tpc_meta;
  P.tdep_ptr = rb.tdep_ptr, P.tdep_idx = rb.tdep_idx;
  P.n_blocks = rb.n_blocks, P.n_rows = A.n_rows, P.n_tiles = rb.n_tiles, P.cap_len = rb.tile_len_max, P.cap_rows = rb.tile_rows_max;
  P.n_hops = n, P.n_slabs = ceil_div(F, 64), P.F = F;
  P.debug_skip = (int)g_options[DSW_OPT_DEBUG].load(std::memory_order_relaxed);

  // shared memory per team: staged rows + panels
  const size_t team_bytes = ((size_t)rb.tile_rows_max * 256 + (size_t)(rb.tile_len_max + DSW_PANEL_PAD) * DSW_TILE_BLOCKS * 20 + 127) & ~(size_t)127;
  const size_t smem = CH_TEAMS * team_bytes + 64 + CH_TEAMS * sizeof(ItemDesc) + 64;
  if (smem > 227 * 1024) return DSW_ERR_UNSUPPORTED;
  P.team_stride = (uint32_t)team_bytes;

  // Samples per group: three planes of a group inside the L2 budget, but a hop pass of at least ~2 items per team.
  const int n_sm = sm_count();
  int64_t budget = g_options[DSW_OPT_CHAIN_L2_BYTES].load(std::memory_order_relaxed);
  if (budget <= 0) budget = (int64_t)64 << 20;
  const int64_t per_sample = 3 * (int64_t)A.n_rows * F * (int64_t)sizeof(float);
  int64_t S = std::max<int64_t>(1, budget / std::max<int64_t>(per_sample, 1));
  const int64_t min_pass = g_options[DSW_OPT_CHAIN_MIN_PASS].load(std::memory_order_relaxed) > 0
                               ? g_options[DSW_OPT_CHAIN_MIN_PASS].load(std::memory_order_relaxed)
                               : (int64_t)2 * n_sm * CH_TEAMS;
  S = std::max<int64_t>(S, ceil_div64(min_pass, (int64_t)rb.n_tiles * P.n_slabs));
  S = std::min<int64_t>(S, Bn);
  P.S = (int32_t)S;
  P.n_full_groups = Bn / P.S;
  P.S_last = Bn - P.n_full_groups * P.S;
  const int64_t items_full = (int64_t)n * rb.n_tiles * P.S * P.n_slabs;
  const int64_t total = items_full * P.n_full_groups + (int64_t)n * rb.n_tiles * P.S_last * P.n_slabs;
  if (total >= ((int64_t)1 << 30)) return DSW_ERR_UNSUPPORTED;
  P.items_full = (int32_t)items_full;
  P.total_items = (int32_t)total;

  const uint32_t k = rb.chain_ring->fetch_add(1, std::memory_order_relaxed);
  P.sync = rb.chain_sync + (size_t)(k % DSW_CHAIN_SETS) * (DSW_CHAIN_HDR + (size_t)rb.chain_flag_cap);
  P.n_ctas = (int32_t)std::min<int64_t>(n_sm, ceil_div64(total, CH_TEAMS));

  ChainArgs args;
  ChainMaps maps;
  for (int j = 0; j < n; ++j) {
    ChainHop c = hops[j];
    // sample window [B0, B0 + Bn)
    c.X += B0 * c.x_sB, c.O += B0 * c.o_sB;
    if (c.Z) c.Z += B0 * c.z_sB;
    if (c.G) c.G += B0 * c.g_sB;
    c.dep = j > 0 ? 1 : 0;
    args.h[j] = c;
    const uint64_t dims[3] = {(uint64_t)F, (uint64_t)A.n_cols, (uint64_t)Bn};
    const uint64_t strides[2] = {(uint64_t)c.x_sV * 4, (uint64_t)std::max<int64_t>(c.x_sB, 1) * 4};
    for (int q = 0; q < 8; ++q) {
      const uint32_t box[3] = {64u, 1u << q, 1u};
      if ((1 << q) > A.n_cols) {
        maps.m[j][q] = maps.m[j][q - 1];
      } else if (!encode_f32_map(&maps.m[j][q], c.X, 3, dims, strides, box)) {
        return DSW_ERR_UNSUPPORTED;
      }
    }
  }
  for (int j = n; j < DSW_CHAIN_MAX_HOPS; ++j) {
    args.h[j] = ChainHop{};
    for (int q = 0; q < 8; ++q) maps.m[j][q] = maps.m[0][q];
  }

  static std::atomic<int> attr_dev_mask[2] = {{0}, {0}};  // per device ordinal (0..63): attribute applied
  int dev = 0;
  DSW_CUDA_TRY(cudaGetDevice(&dev));
  {
    std::atomic<int>& m = attr_dev_mask[(dev >> 5) & 1];
    if (!(m.load(std::memory_order_acquire) & (1 << (dev & 31)))) {
      DSW_CUDA_TRY(cudaFuncSetAttribute(hop_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      m.fetch_or(1 << (dev & 31), std::memory_order_release);
    }
  }
  DSW_CUDA_TRY(launch_pdl(hop_chain_kernel, dim3(P.n_ctas), dim3(CH_THREADS), smem, st, pdl_enabled(), P, args, maps));
  return check_launch();
}

int launch_hop_chain(const dsw_csr& A, const dsw_rb& rb, const ChainHop* hops, int n, int32_t B, int32_t F, cudaStream_t st) {
  if (n <= 0) return DSW_OK;
  const int min_hops = (int)std::max<int64_t>(1, g_options[DSW_OPT_CHAIN_MIN_HOPS].load(std::memory_order_relaxed));
  if (n >= min_hops && chain_ok(A, rb, hops, n, B, F)) {
    // sample windows that fit one set of flags; hop windows of at most DSW_CHAIN_MAX_HOPS (the launch boundary orders them)
    const int64_t per_sample = (int64_t)ceil_div(F, 64) * rb.n_tiles;
    const int32_t Bw = (int32_t)std::min<int64_t>(B, std::max<int64_t>(1, rb.chain_flag_cap / per_sample));
    if (per_sample <= rb.chain_flag_cap) {
      int rc = DSW_OK;
      for (int j0 = 0; j0 < n && rc == DSW_OK; j0 += DSW_CHAIN_MAX_HOPS)
        for (int32_t b0 = 0; b0 < B && rc == DSW_OK; b0 += Bw)
          rc = launch_chain_fused(A, rb, hops + j0, std::min(DSW_CHAIN_MAX_HOPS, n - j0), b0, std::min(Bw, B - b0), F, st);
      if (rc != DSW_ERR_UNSUPPORTED) return rc;
    }
  }
  for (int j = 0; j < n; ++j) {
    HopArgs a;
    const ChainHop& c = hops[j];
    a.X = c.X, a.x_sB = c.x_sB, a.x_sV = c.x_sV;
    a.Z = c.Z, a.z_sB = c.z_sB, a.z_sV = c.z_sV;
    a.G = c.G, a.g_sB = c.g_sB, a.g_sV = c.g_sV;
    a.O = c.O, a.o_sB = c.o_sB, a.o_sV = c.o_sV;
    a.alpha = c.alpha, a.beta = c.beta, a.B = B, a.F = F, a.act = c.act;
    DSW_TRY(launch_hop(A, rb, a, st));
  }
  return DSW_OK;
}

}  // namespace dsw

extern "C" int dsw_debug_chain_counters(uint64_t* out8, int reset) {
  if (!out8) return DSW_ERR_BAD_ARGUMENT;
  unsigned long long h[8];
  DSW_CUDA_TRY(cudaMemcpyFromSymbol(h, dsw::g_chain_prof, sizeof(h)));
  for (int i = 0; i < 8; ++i) out8[i] = h[i];
  if (reset) {
    unsigned long long z[8] = {};
    DSW_CUDA_TRY(cudaMemcpyToSymbol(dsw::g_chain_prof, z, sizeof(z)));
  }
  return DSW_OK;
}
