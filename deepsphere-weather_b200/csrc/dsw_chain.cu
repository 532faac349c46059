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
//                 grid (one CTA per SM, CH_TEAMS independent teams per CTA)
//   group size  = as many samples as keep three planes of the group (gather source, the k-2 term, the output)
//                 inside the L2 budget, but enough for a hop pass to hold more items than are in flight.
//
// Dependencies instead of grid barriers: item (b, j, t) gathers rows of O_{j-1} that lie in a fixed, small set of
// tiles deps(t) (plan: tdep_fix, own tile included).  Every finished item publishes
// flag[b][slab][t] = epoch * 16 + j + 1 behind a gpu-scope release; an item's flags are polled before its transfers
// are issued.  Claims are handed out in order and an item only ever waits for smaller claims, so the grid cannot
// deadlock whatever its size.  The epoch lives in device memory and is bumped by the last CTA to leave, which also
// zeroes the claim counter: no memset launches, and captured CUDA graphs replay correctly.
//
// Warp roles (512 threads, registers rebalanced with setmaxnreg):
//   warps 8-15  four compute teams of two warps (the issue arbiter favours high warp ids: the spinning control warps
//               never take an issue slot a compute warp could use).  4 lanes own a row-block of 4 rows x 64 channels (64 fp32
//               accumulators per lane); every 16-byte shared-memory read feeds 4 packed FMAs (FFMA2) — the
//               register-tiled loop of hop_team_kernel (dsw_spmm.cu).  They only ever wait on mbarriers.
//   warps 0-3   one issuer warp per team, running one item ahead of it: claim, decode, the tile's metadata (one
//               round trip: fixed-stride plan tables), dependency flags (a second one); the moment the team's
//               entry loop ends it posts the descriptor (`ready`: the team starts its Z / G loads) and issues the transfers —
//               the tile's weight / offset panels (two bulk copies) and the distinct source rows the tile gathers (one
//               tensor-map box per run of consecutive rows), in two phases: the run that holds the tile's own rows lands on
//               `full` with the panels, the other rows on `full2`; the entry loop starts on the steps that need only the former.
//   warp 4      publisher: collects the items whose stores have been issued and releases their flags behind ONE
//               gpu-scope fence per round (a MEMBAR.ALL.GPU costs ~7k cycles on a busy SM: it must not sit on a
//               compute or issuer warp's path).
#include <algorithm>

#include "dsw_internal.cuh"
#include "dsw_tmap.cuh"

namespace dsw {

constexpr int CH_TEAMS = 4;
constexpr int CH_TEAM_THREADS = DSW_TILE_BLOCKS * 4;  // 4 lanes per row-block: two compute warps per team
constexpr int CH_THREADS = 512;                       // 8 compute warps, 4 issuer warps, publisher + 3 idle warps
constexpr int CH_DESC_RING = 4;                       // item descriptors per team
constexpr int CH_REGS_COMPUTE = 200, CH_REGS_ISSUER = 64, CH_REGS_PUBLISHER = 40;

struct ChainPlan {
  const int32_t* blkptr;
  const int32_t* tp_ptr;
  const float4* tp_val;
  const uint32_t* tp_off;
  const int4* tile_meta;
  const int2* tpc_fix;
  const int32_t* tdep_fix;
  int32_t pieces_stride, deps_stride;
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

__device__ unsigned long long g_chain_prof[16];  // [0..7] compute teams, [8..15] control warps

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
// Dependency flags are polled with a strong relaxed load.  What is read once a flag is seen — the source rows (TMA)
// and the Z / G rows (ld.global.cg) — is fetched from L2, the point of coherence, where the producer's gpu-scope
// release (MEMBAR.ALL.GPU before the flag store) has put it; no L1 line is involved, so the L1 invalidation an
// acquire load would cost the whole SM on every poll (CCTL.IVALL) buys nothing here.
__device__ __forceinline__ int32_t ld_flag(const int32_t* p) {
  int32_t v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
// L2 eviction hints: a Z row is read for the last time by this launch, and the last hop's output is not read by it at all —
// both leave the L2 first, which keeps the planes the next hop pass still needs resident.
__device__ __forceinline__ uint64_t l2_policy(bool evict_first) {
  uint64_t p;
  if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldcg4_hint(const float* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol)
               : "memory");
  return v;
}
__device__ __forceinline__ void st4_hint(float* p, const float4& v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
               : "memory");
}
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

__device__ __forceinline__ void fma_step2(float4 (&acc)[4][4], const float4& w, const float4 (&x)[2]) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    fma4(acc[0][j], w.x, x[j]);
    fma4(acc[1][j], w.y, x[j]);
    fma4(acc[2][j], w.z, x[j]);
    fma4(acc[3][j], w.w, x[j]);
  }
}

// What a team needs to know about its current item (written by its issuer warp, read after the `ready` barrier).
struct ItemDesc {
  int32_t idx;   // claim index (< 0: nothing left, leave)
  int32_t hop, tile, b, slab;
  int32_t wlen[2];  // entry-loop trip count of the team's first / second warp (plan: tile_meta)
  int32_t split;    // plan: tile_split (pieces of group A | rows of A << 8 | steps on A only, warp 0 << 16 | warp 1 << 24)
};

// mbarriers of one team
struct TeamBars {
  uint64_t ready;  // issuer -> team: descriptor written, dependencies met, transfers about to be issued (count 1)
  uint64_t full;   // transfers -> team: the panels and the rows of group A have landed (count 1 + tx bytes)
  uint64_t full2;  // transfers -> team: the other rows have landed (count 1 + tx bytes)
  uint64_t empty;  // team -> issuer: the entry loop is over, the buffers may be overwritten (count 64: every lane)
};

// plain words shared between the roles of one team
struct TeamWords {
  uint32_t done[2];    // per compute warp: items whose stores have been issued (release, cta)
  uint32_t published;  // items whose flag is out (publisher)
  int32_t total;       // items the issuer handed to the team, -1 while it is still claiming
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void red_release_cta(uint32_t* p, bool relaxed_for_timing_only = false) {
  if (relaxed_for_timing_only)
    asm volatile("red.relaxed.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(p)) : "memory");
  else
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(p)) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_volatile_s(const void* p) {
  uint32_t v;
  asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_s(void* p, uint32_t v) {
  asm volatile("st.volatile.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

}  // namespace

// MASK: the hops may carry a ReLU mask operand (ChainHop::M).  A separate instantiation: the plain kernel's code
// generation is sensitive to anything added to its store loop (measured: 808 -> 888 us at the metric shape with the
// masked store compiled in, although no hop used it).
template <bool MASK>
__global__ void __launch_bounds__(CH_THREADS, 1)
    hop_chain_kernel(const ChainPlan P, const __grid_constant__ ChainArgs A, const __grid_constant__ ChainMaps maps) {
  extern __shared__ __align__(1024) uint8_t ch_smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  // shared memory: [team 0: rows | weights | offsets] ... [team 3] | mbarriers | words | item descriptors | epoch
  uint8_t* tail = ch_smem + (size_t)CH_TEAMS * P.team_stride;
  TeamBars* s_bars = reinterpret_cast<TeamBars*>(tail);
  TeamWords* s_words = reinterpret_cast<TeamWords*>(tail + CH_TEAMS * sizeof(TeamBars));
  ItemDesc* s_item = reinterpret_cast<ItemDesc*>(tail + CH_TEAMS * (sizeof(TeamBars) + sizeof(TeamWords)));
  int32_t* s_epoch = reinterpret_cast<int32_t*>(tail + CH_TEAMS * (sizeof(TeamBars) + sizeof(TeamWords) + CH_DESC_RING * sizeof(ItemDesc)));

  pdl_trigger();
  if (tid < CH_TEAMS) {
    mbar_init(smem_u32(&s_bars[tid].ready), 1);
    mbar_init(smem_u32(&s_bars[tid].full), 1);
    mbar_init(smem_u32(&s_bars[tid].full2), 1);
    mbar_init(smem_u32(&s_bars[tid].empty), CH_TEAM_THREADS);
    s_words[tid].done[0] = 0, s_words[tid].done[1] = 0, s_words[tid].published = 0, s_words[tid].total = -1;
  }
  if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  // everything the previous kernel of the stream wrote (operands, and — if the ring of sets has wrapped — this set's
  // sync words) is visible from here on
  pdl_wait();
  if (tid == 0) *s_epoch = *reinterpret_cast<volatile int32_t*>(P.sync + 2);
  __syncthreads();
  const int32_t base = (*s_epoch + 1) * 16;  // flag value of "hop j done" = base + j + 1
  int32_t* const claim = P.sync;
  int32_t* const flags = P.sync + DSW_CHAIN_HDR;

  // Warp roles by warp id: 0-3 issuers, 4 publisher (5-7 idle), 8-15 compute.  The issue arbiter favours high warp
  // ids, so the spinning control warps never take an issue slot a compute warp could use.
  if (warp >= 4 && warp < 8) {
    // ========================================= publisher =========================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CH_REGS_PUBLISHER));
    if (warp == 4) {
      uint32_t pn = 0;  // lane t < CH_TEAMS: items of team t published so far
      bool finished = lane >= CH_TEAMS;
      while (true) {
        uint32_t avail = 0;
        if (!finished) {
          avail = min(ld_acquire_cta(&s_words[lane].done[0]), ld_acquire_cta(&s_words[lane].done[1]));
          const int32_t total = (int32_t)ld_volatile_s(&s_words[lane].total);
          if (total >= 0 && pn == (uint32_t)total) finished = true;
        }
        const bool mine = !finished && avail > pn;
        if (__any_sync(0xffffffffu, mine)) {
          // one gpu-scope release for every item collected this round: the teams' stores (ordered before `done` at cta
          // scope) reach L2 before any of the flags below does
          __threadfence();
          if (mine) {
            for (; pn < avail; ++pn) {
              const ItemDesc* d = &s_item[lane * CH_DESC_RING + (pn % CH_DESC_RING)];
              const int32_t hop = (int32_t)ld_volatile_s(&d->hop), tile = (int32_t)ld_volatile_s(&d->tile);
              const int32_t b = (int32_t)ld_volatile_s(&d->b), slab = (int32_t)ld_volatile_s(&d->slab);
              asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(flags + ((int64_t)b * P.n_slabs + slab) * P.n_tiles + tile),
                           "r"(base + hop + 1)
                           : "memory");
            }
            st_volatile_s(&s_words[lane].published, pn);
          }
        } else if (__all_sync(0xffffffffu, finished)) {
          break;
        } else {
          __nanosleep(400);
        }
      }
    }
  } else if (warp < 4) {
    // ========================================== issuer ===========================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CH_REGS_ISSUER));
    const int team = warp;
    uint8_t* xs = ch_smem + (size_t)team * P.team_stride;
    const uint32_t xs_u32 = smem_u32(xs);
    const uint32_t sval_u32 = xs_u32 + (uint32_t)P.cap_rows * 256u;
    const uint32_t soff_u32 = sval_u32 + (uint32_t)(P.cap_len + DSW_PANEL_PAD) * DSW_TILE_BLOCKS * 16u;
    const uint32_t bar_ready = smem_u32(&s_bars[team].ready), bar_full = smem_u32(&s_bars[team].full);
    const uint32_t bar_full2 = smem_u32(&s_bars[team].full2);
    const uint32_t bar_empty = smem_u32(&s_bars[team].empty);
    const bool cprof = (P.debug_skip & 4) && (lane == 0);

    int32_t idx = 0;
    if (lane == 0) idx = atomicAdd(claim, 1);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    for (uint32_t n = 0;; ++n) {
      long long k0 = 0, k1 = 0, k2 = 0, k3 = 0;
      if (cprof) k0 = clock64();
      if (idx >= P.total_items) {
        // nothing left: the team leaves after its current item, the publisher after the team's last flag
        if (n > 0) mbar_wait(bar_empty, (n - 1) & 1u);
        while (n >= CH_DESC_RING && ld_volatile_s(&s_words[team].published) + CH_DESC_RING <= n) __nanosleep(40);
        if (lane == 0) {
          st_volatile_s(&s_item[team * CH_DESC_RING + (n % CH_DESC_RING)].idx, (uint32_t)-1);
          st_volatile_s(&s_words[team].total, n);
          mbar_arrive(bar_ready);
        }
        break;
      }
      int32_t idx_next = 0;
      const bool claim_ahead = (P.debug_skip & 32) != 0;  // measured slower: the deeper claim window makes dependencies late
      if (lane == 0 && claim_ahead) idx_next = atomicAdd(claim, 1);  // consumed at the top of the next round
      // ---- decode item n and fetch its metadata (the team is still busy with item n - 1) ----
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
      // one round trip: tile record, this lane's pieces, this lane's dependency
      const int4 tm = __ldg(P.tile_meta + 2 * tile);  // {panel step offset, steps, source rows, pieces}
      const int4 tm2 = __ldg(P.tile_meta + 2 * tile + 1);  // {deps, trip count of warp 0, of warp 1, split} (same 32-byte sector)
      const int32_t split = (P.debug_skip & 4096) ? 0 : tm2.w;
      const int n_a = split & 0xff;
      const uint32_t rows_a = ((uint32_t)split >> 8) & 0xffu;
      int2 pc0 = make_int2(-1, 0), pc1 = make_int2(-1, 0);
      if (lane < P.pieces_stride) pc0 = __ldg(P.tpc_fix + (size_t)tile * P.pieces_stride + lane);
      if (lane + 32 < P.pieces_stride) pc1 = __ldg(P.tpc_fix + (size_t)tile * P.pieces_stride + lane + 32);
      const bool dep = A.h[hop].dep != 0;
      int32_t dtile = -1;
      if (dep && lane < P.deps_stride) dtile = __ldg(P.tdep_fix + (size_t)tile * P.deps_stride + lane);
      const int32_t need = base + hop;  // hop - 1 finished
      const int32_t* fdep = dtile >= 0 ? flags + ((int64_t)b * P.n_slabs + slab) * P.n_tiles + dtile : nullptr;
      // a second one: an early look at the dependencies (they are a whole hop pass old and almost always met by now)
      bool dep_ok = __all_sync(0xffffffffu, fdep == nullptr || ld_flag(fdep) - need >= 0);
      if (cprof) k1 = clock64();

      // ---- the team's buffers are free once its entry loop of item n - 1 is over ----
      if (n > 0) mbar_wait(bar_empty, (n - 1) & 1u);
      if (cprof) k2 = clock64();
      bool late = false;
      if (!dep_ok) dep_ok = __all_sync(0xffffffffu, fdep == nullptr || ld_flag(fdep) - need >= 0);
      if (!dep_ok) {
        // A blocking wait must not start before every earlier item of this team is published (item n may depend on
        // them); the team pays its `done` arrival at once when it finds no next item.
        late = true;
        while (ld_volatile_s(&s_words[team].published) < n) __nanosleep(40);
        bool ok = fdep == nullptr || ld_flag(fdep) - need >= 0;
        while (!ok) {
          __nanosleep(64);
          ok = ld_flag(fdep) - need >= 0;
        }
        __syncwarp();
      }
      // the descriptor slot is free once the publisher is done with item n - CH_DESC_RING
      while (n >= CH_DESC_RING && ld_volatile_s(&s_words[team].published) + CH_DESC_RING <= n) __nanosleep(40);
      // the team read the buffers through the generic proxy; the transfers write them through the async proxy
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (lane == 0) {
        ItemDesc d;
        d.idx = idx, d.hop = hop, d.tile = tile, d.b = b, d.slab = slab, d.wlen[0] = tm2.y, d.wlen[1] = tm2.z, d.split = split;
        s_item[team * CH_DESC_RING + (n % CH_DESC_RING)] = d;
        // two phases: the panels + the rows of group A (the run that holds the tile's own rows: 2-3 boxes), then the rest
        mbar_expect_tx(bar_full, rows_a * 256u + (uint32_t)tm.y * (DSW_TILE_BLOCKS * 20u));
        mbar_expect_tx(bar_full2, ((uint32_t)tm.z - rows_a) * 256u);
        // `ready` goes out BEFORE the ~24 transfer instructions (~1.4 k cycles of issue): the team reads the descriptor and
        // sends its Z / G loads while the boxes are still being issued, then waits on `full`
        mbar_arrive(bar_ready);
        bulk_g2s(sval_u32, P.tp_val + (size_t)tm.x * DSW_TILE_BLOCKS, (uint32_t)tm.y * (DSW_TILE_BLOCKS * 16u), bar_full);
        bulk_g2s(soff_u32, P.tp_off + (size_t)tm.x * DSW_TILE_BLOCKS, (uint32_t)tm.y * (DSW_TILE_BLOCKS * 4u), bar_full);
      }
      __syncwarp();
      const CUtensorMap* mp = &maps.m[hop][0];
      // (the pieces of group A are the first n_a of the tile's list: lanes 0 .. n_a - 1 issue them first)
      if (pc0.x != -1)
        tma_load_3d(xs_u32 + ((uint32_t)pc0.x >> 8) * 256u, mp + (pc0.x & 7), slab * 64, pc0.y, b, lane < n_a ? bar_full : bar_full2);
      if (pc1.x != -1)
        tma_load_3d(xs_u32 + ((uint32_t)pc1.x >> 8) * 256u, mp + (pc1.x & 7), slab * 64, pc1.y, b, lane + 32 < n_a ? bar_full : bar_full2);
      for (int i = lane + 64; i < tm.w; i += 32) {
        const int2 pc = __ldg(P.tpc_fix + (size_t)tile * P.pieces_stride + i);
        tma_load_3d(xs_u32 + ((uint32_t)pc.x >> 8) * 256u, mp + (pc.x & 7), slab * 64, pc.y, b, i < n_a ? bar_full : bar_full2);
      }
      __syncwarp();
      if (cprof && n > 0) {
        k3 = clock64();
        atomicAdd(&g_chain_prof[8], (unsigned long long)(k1 - k0));   // claim, decode, metadata, early dependency look
        atomicAdd(&g_chain_prof[9], (unsigned long long)(k2 - k1));   // waiting for the team's entry loop to end
        atomicAdd(&g_chain_prof[10], (unsigned long long)(k3 - k2));  // dependency re-check (+ blocking wait) + issue
        atomicAdd(&g_chain_prof[14], late ? 1ull : 0ull);             // items whose dependencies were not met in time
        atomicAdd(&g_chain_prof[15], 1ull);
      }
      if (lane == 0 && !claim_ahead) idx_next = atomicAdd(claim, 1);
      idx = __shfl_sync(0xffffffffu, idx_next, 0);
    }
  } else {
    // ======================================= compute team =======================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CH_REGS_COMPUTE));
    const int ctid = tid - 256;  // compute thread id
    const int team = ctid / CH_TEAM_THREADS;
    const int tt = ctid - team * CH_TEAM_THREADS;
    const uint8_t* xs = ch_smem + (size_t)team * P.team_stride;
    const float4* s_val = reinterpret_cast<const float4*>(xs + (size_t)P.cap_rows * 256);
    const uint32_t* s_off =
        reinterpret_cast<const uint32_t*>(xs + (size_t)P.cap_rows * 256 + (size_t)(P.cap_len + DSW_PANEL_PAD) * DSW_TILE_BLOCKS * 16);
    const uint32_t bar_ready = smem_u32(&s_bars[team].ready), bar_full = smem_u32(&s_bars[team].full);
    const uint32_t bar_full2 = smem_u32(&s_bars[team].full2);
    const uint32_t bar_empty = smem_u32(&s_bars[team].empty);
    uint32_t* const done = &s_words[team].done[tt >> 5];
    const int slot = tt >> 2, lq = tt & 3, par = slot & 1;
    const uint32_t cA = ((uint32_t)(lq * 16) ^ (uint32_t)(par * 64));
    const uint32_t cB = ((uint32_t)(lq * 16 + 64) ^ (uint32_t)(par * 64));
    const uint8_t* xA = xs + cA;
    const uint8_t* xB = xs + cB;
    const uint32_t* po = s_off + slot;
    const float4* pw = s_val + slot;
    int ch[4];
    ch[0] = (int)(cA >> 2), ch[1] = (int)(cB >> 2), ch[2] = ch[0] + 32, ch[3] = ch[1] + 32;

    const uint64_t pol_normal = l2_policy(false);
    const uint64_t pol_first = l2_policy(true);

    for (uint32_t n = 0;; ++n) {
      const bool prof = (P.debug_skip & 4) && (tt == 0);
      long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
      if (prof) c0 = clock64();
      // The `done` arrival of the previous item orders its stores (a CTA-scope memory barrier that waits for them).
      // If the next item is already there, its Z / G loads go out first and the barrier's wait overlaps the transfers;
      // otherwise there is nothing better to do than to pay it now (and the issuer may need it before it can wait for
      // the next item's dependencies).
      const bool early = n > 0 && __shfl_sync(0xffffffffu, (int)mbar_test(bar_ready, n & 1u), 0) != 0;
      if (n > 0 && !early) {
        __syncwarp();
        if (lane == 0) red_release_cta(done, (P.debug_skip & 64) != 0);
      }
      mbar_wait(bar_ready, n & 1u);
      const ItemDesc d = s_item[team * CH_DESC_RING + (n % CH_DESC_RING)];
      if (d.idx < 0) {
        if (early) {
          __syncwarp();
          if (lane == 0) red_release_cta(done, (P.debug_skip & 64) != 0);
        }
        break;
      }
      const ChainHop& H = A.h[d.hop];
      const int blk = d.tile * DSW_TILE_BLOCKS + slot;
      const bool active = blk < P.n_blocks;
      const int slab_f = min(64, P.F - d.slab * 64);
      if (prof) c1 = clock64();

      // Accumulators start at (beta * Z + G) / alpha (alpha is 1 or 2: exact); the loads land while the tile does.
      // Whole row-blocks of a full 64-channel slab (every item but those of a ragged last tile / last slab) take a
      // straight-line path: one base address per operand, no per-element range checks — the checked form costs ~600
      // instructions per item and warp before the entry loop and as many after it, about as many as the loop itself.
      const bool fast = slab_f == 64 && __all_sync(0xffffffffu, active && blk * 4 + 3 < P.n_rows) && !(P.debug_skip & (8 | 16));
      const float h_alpha = H.alpha;
      float4 acc[4][4];
      if (fast) {
        const float inv_alpha = h_alpha == 2.f ? 0.5f : 1.f;  // (the launcher admits alpha 1 and 2 only)
        const float zs = H.beta * inv_alpha;
        const float* const hZ = H.Z;
        const float* const hG = H.G;
        const int64_t row0 = (int64_t)blk * 4;
        if (hZ != nullptr) {
          const int64_t sv = H.z_sV;
          const float* zp = hZ + d.b * H.z_sB + row0 * sv + d.slab * 64;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[r][j] = ldcg4_hint(zp + ch[j], pol_first);
            zp += sv;
          }
        }
        if (hG != nullptr) {
          float4 g[4][4];
          const int64_t sv = H.g_sV;
          const float* gp = hG + d.b * H.g_sB + row0 * sv + d.slab * 64;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int j = 0; j < 4; ++j) g[r][j] = ldcg4_hint(gp + ch[j], pol_first);
            gp += sv;
          }
          if (early) {
            __syncwarp();
            if (lane == 0) red_release_cta(done, (P.debug_skip & 64) != 0);
          }
          if (hZ != nullptr) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                // (same operations as the checked path: round(z * zs), then one fused multiply-add)
                const float4 z = make_float4(acc[r][j].x * zs, acc[r][j].y * zs, acc[r][j].z * zs, acc[r][j].w * zs);
                acc[r][j] = make_float4(fmaf(g[r][j].x, inv_alpha, z.x), fmaf(g[r][j].y, inv_alpha, z.y),
                                        fmaf(g[r][j].z, inv_alpha, z.z), fmaf(g[r][j].w, inv_alpha, z.w));
              }
          } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                acc[r][j] = make_float4(fmaf(g[r][j].x, inv_alpha, 0.f), fmaf(g[r][j].y, inv_alpha, 0.f),
                                        fmaf(g[r][j].z, inv_alpha, 0.f), fmaf(g[r][j].w, inv_alpha, 0.f));
          }
        } else {
          if (early) {
            __syncwarp();
            if (lane == 0) red_release_cta(done, (P.debug_skip & 64) != 0);
          }
          if (hZ != nullptr) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                acc[r][j] = make_float4(acc[r][j].x * zs, acc[r][j].y * zs, acc[r][j].z * zs, acc[r][j].w * zs);
          } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      } else {
        const float inv_alpha = 1.f / h_alpha;
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
            if (H.Z != nullptr && ok[r][j] && !(P.debug_skip & 16))
              acc[r][j] = ldcg4(H.Z + d.b * H.z_sB + (int64_t)(blk * 4 + r) * H.z_sV + d.slab * 64 + ch[j]);
          }
        if (early && H.G == nullptr) {
          __syncwarp();
          if (lane == 0) red_release_cta(done, (P.debug_skip & 64) != 0);
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
          if (early) {
            __syncwarp();
            if (lane == 0) red_release_cta(done, (P.debug_skip & 64) != 0);
          }
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              acc[r][j] = make_float4(fmaf(g[r][j].x, inv_alpha, acc[r][j].x), fmaf(g[r][j].y, inv_alpha, acc[r][j].y),
                                      fmaf(g[r][j].z, inv_alpha, acc[r][j].z), fmaf(g[r][j].w, inv_alpha, acc[r][j].w));
        }
      }
      const int wlen = (tt >> 5) ? d.wlen[1] : d.wlen[0];  // (no global load on the team's path: it used to delay the Z / G loads by an L2 round trip)
      // entry steps [0, wsplit) gather from group A only (the run of rows that holds the tile's own rows: it lands first, with
      // the panels); the loop runs them while the other boxes are still being issued / in flight, then waits for `full2`
      const int wsplit = min(wlen, (int)(((uint32_t)d.split >> ((tt >> 5) ? 24 : 16)) & 0xffu));
      const int n_steps = (P.debug_skip & 2) ? 0 : wlen;

      if (slab_f > 32) {
        auto load_x = [&](uint32_t o, float4(&x)[4]) {
          x[0] = *reinterpret_cast<const float4*>(xA + o);
          x[1] = *reinterpret_cast<const float4*>(xB + o);
          x[2] = *reinterpret_cast<const float4*>(xA + o + 128);
          x[3] = *reinterpret_cast<const float4*>(xB + o + 128);
        };
#pragma unroll 1
        for (int seg = 0; seg < 2; ++seg) {
          mbar_wait(seg ? bar_full2 : bar_full, n & 1u);
          if (prof && seg == 0) c2 = clock64();
          const int u0 = seg ? wsplit : 0, u1 = seg ? n_steps : min(wsplit, n_steps);
          if (u0 >= u1) continue;
          // software pipeline: offsets two steps ahead, weights / values one step ahead (the panels end in zero steps; what
          // the last trip of the first segment reads ahead may not have landed yet and is discarded)
          float4 x0[4], x1[4], w0, w1;
          uint32_t o1, o2;
          w0 = pw[u0 * DSW_TILE_BLOCKS];
          load_x(po[u0 * DSW_TILE_BLOCKS], x0);
          o1 = po[(u0 + 1) * DSW_TILE_BLOCKS];
#pragma unroll 1
          for (int u = u0; u < u1; u += 2) {
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
      } else {
        // a slab of at most 32 channels (24-channel first layer, narrow last slabs): accumulator columns 2 and 3 (channels
        // 32-63) are never stored — half the shared-memory reads and FMAs
        auto load_x2 = [&](uint32_t o, float4(&x)[2]) {
          x[0] = *reinterpret_cast<const float4*>(xA + o);
          x[1] = *reinterpret_cast<const float4*>(xB + o);
        };
        mbar_wait(bar_full, n & 1u);
        mbar_wait(bar_full2, n & 1u);
        if (prof) c2 = clock64();
        float4 x0[2], x1[2], w0, w1;
        uint32_t o1, o2;
        w0 = pw[0];
        load_x2(po[0], x0);
        o1 = po[DSW_TILE_BLOCKS];
#pragma unroll 1
        for (int u = 0; u < n_steps; u += 2) {
          w1 = pw[(u + 1) * DSW_TILE_BLOCKS];
          load_x2(o1, x1);
          o2 = po[(u + 2) * DSW_TILE_BLOCKS];
          fma_step2(acc, w0, x0);
          w0 = pw[(u + 2) * DSW_TILE_BLOCKS];
          load_x2(o2, x0);
          o1 = po[(u + 3) * DSW_TILE_BLOCKS];
          fma_step2(acc, w1, x1);
        }
      }
      // this lane is done with the staged rows and panels (every lane arrives itself: its own shared-memory reads are
      // then ordered before the issuer's transfers without leaning on a warp-level sync)
      mbar_arrive(bar_empty);
      if (prof) c3 = clock64();
      // ---- epilogue: O = alpha * acc ----
      if (fast) {
        const int64_t sv = H.o_sV;
        float* op = H.O + d.b * H.o_sB + (int64_t)blk * 4 * sv + d.slab * 64;
        const float* mp = nullptr;
        int64_t msv = 0;
        if (MASK && H.M != nullptr) msv = H.m_sV, mp = H.M + d.b * H.m_sB + (int64_t)blk * 4 * msv + d.slab * 64;
        const bool relu = H.act != 0;
        const uint64_t pol_out = d.hop == P.n_hops - 1 ? pol_first : pol_normal;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 o = make_float4(h_alpha * acc[r][j].x, h_alpha * acc[r][j].y, h_alpha * acc[r][j].z, h_alpha * acc[r][j].w);
            if (relu) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
            if (MASK && mp != nullptr) {  // ReLU mask of a gradient by the ReLU's output (indexed like the output rows)
              const float4 m = ldcg4(mp + ch[j]);
              o.x = m.x > 0.f ? o.x : 0.f, o.y = m.y > 0.f ? o.y : 0.f, o.z = m.z > 0.f ? o.z : 0.f, o.w = m.w > 0.f ? o.w : 0.f;
            }
            st4_hint(op + ch[j], o, pol_out);
          }
          op += sv;
          if (MASK) mp += msv;
        }
      } else if (active && !(P.debug_skip & 8)) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int row = blk * 4 + r;
          if (row >= P.n_rows) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (ch[j] >= slab_f) continue;
            float4 o = make_float4(H.alpha * acc[r][j].x, H.alpha * acc[r][j].y, H.alpha * acc[r][j].z, H.alpha * acc[r][j].w);
            if (H.act) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
            if (MASK && H.M != nullptr) {  // ReLU mask of a gradient by the ReLU's output (indexed like the output rows)
              const float4 m = ldcg4(H.M + d.b * H.m_sB + (int64_t)row * H.m_sV + d.slab * 64 + ch[j]);
              o.x = m.x > 0.f ? o.x : 0.f, o.y = m.y > 0.f ? o.y : 0.f, o.z = m.z > 0.f ? o.z : 0.f, o.w = m.w > 0.f ? o.w : 0.f;
            }
            *reinterpret_cast<float4*>(H.O + d.b * H.o_sB + (int64_t)row * H.o_sV + d.slab * 64 + ch[j]) = o;
          }
        }
      }
      // (the `done` arrival follows at the top of the next round)
      if (prof) {
        c4 = clock64();
        atomicAdd(&g_chain_prof[0], (unsigned long long)(c1 - c0));  // wait for the next item (issuer, dependencies)
        atomicAdd(&g_chain_prof[1], (unsigned long long)(c2 - c1));  // Z / G loads + wait for the transfers
        atomicAdd(&g_chain_prof[2], (unsigned long long)(c3 - c2));  // entry loop
        atomicAdd(&g_chain_prof[3], (unsigned long long)(c4 - c3));  // stores
        atomicAdd(&g_chain_prof[4], 1ull);
      }
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
  // Measured (tools/bench_layers.py, tools/bench_chain.py): the fused kernel wins where a hop pass has at least one tile per
  // SM (nside >= 32: 226 vs 260 us at V 12288 / F 64, 802 vs 910 us at nside 64); on the coarse U-Net levels the
  // per-tile launches, which stage a tile's panels once per CTA, are 5-10 % faster.  DSW_OPT_NO_CHAIN = 2 forces it.
  if (g_options[DSW_OPT_NO_CHAIN].load(std::memory_order_relaxed) != 2 && rb.n_tiles < sm_count()) return false;
  if (A.n_rows != A.n_cols || F % 4 || B <= 0 || B > 65535) return false;
  const int64_t small_f_opt = g_options[DSW_OPT_HOP_SMALL_F].load(std::memory_order_relaxed);
  if (F <= (small_f_opt > 0 ? (int)small_f_opt : 8)) return false;
  if (rb.tile_pieces_max <= 0 || rb.tile_pieces_max > rb.tile_rows_max) return false;
  for (int j = 0; j < n; ++j) {
    const ChainHop& c = h[j];
    if (!c.X || !c.O || !aligned16(c.X) || !aligned16(c.O) || (c.x_sB | c.x_sV | c.o_sB | c.o_sV) % 4 || c.x_sV < F) return false;
    if (c.Z && (!aligned16(c.Z) || (c.z_sB | c.z_sV) % 4)) return false;
    if (c.G && (!aligned16(c.G) || (c.g_sB | c.g_sV) % 4)) return false;
    if (c.M && (!aligned16(c.M) || (c.m_sB | c.m_sV) % 4)) return false;
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
  P.tile_meta = rb.tile_meta, P.tpc_fix = rb.tpc_fix, P.tdep_fix = rb.tdep_fix;
  P.pieces_stride = rb.tile_pieces_max, P.deps_stride = rb.tile_deps_max;
  P.n_blocks = rb.n_blocks, P.n_rows = A.n_rows, P.n_tiles = rb.n_tiles, P.cap_len = rb.tile_len_max, P.cap_rows = rb.tile_rows_max;
  P.n_hops = n, P.n_slabs = ceil_div(F, 64), P.F = F;
  P.debug_skip = (int)g_options[DSW_OPT_DEBUG].load(std::memory_order_relaxed);

  // shared memory per team: staged rows + panels
  const size_t team_bytes = ((size_t)rb.tile_rows_max * 256 + (size_t)(rb.tile_len_max + DSW_PANEL_PAD) * DSW_TILE_BLOCKS * 20 + 127) & ~(size_t)127;
  const size_t smem = CH_TEAMS * team_bytes + CH_TEAMS * (sizeof(TeamBars) + sizeof(TeamWords) + CH_DESC_RING * sizeof(ItemDesc)) + 64;
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
    if (c.M) c.M += B0 * c.m_sB;
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

  bool any_mask = false;
  for (int j = 0; j < n; ++j) any_mask = any_mask || hops[j].M != nullptr;
  static PerDeviceOnce attr_set[2];
  if (any_mask) {
    DSW_CUDA_TRY(attr_set[1].max_dynamic_smem(hop_chain_kernel<true>, 227 * 1024));
    DSW_CUDA_TRY(launch_pdl(hop_chain_kernel<true>, dim3(P.n_ctas), dim3(CH_THREADS), smem, st, pdl_enabled(), P, args, maps));
  } else {
    DSW_CUDA_TRY(attr_set[0].max_dynamic_smem(hop_chain_kernel<false>, 227 * 1024));
    DSW_CUDA_TRY(launch_pdl(hop_chain_kernel<false>, dim3(P.n_ctas), dim3(CH_THREADS), smem, st, pdl_enabled(), P, args, maps));
  }
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
    a.M = c.M, a.m_sB = c.m_sB, a.m_sV = c.m_sV;
    DSW_TRY(launch_hop(A, rb, a, st));
  }
  return DSW_OK;
}

}  // namespace dsw

extern "C" int dsw_debug_chain_counters(uint64_t* out16, int reset) {
  if (!out16) return DSW_ERR_BAD_ARGUMENT;
  unsigned long long h[16];
  DSW_CUDA_TRY(cudaMemcpyFromSymbol(h, dsw::g_chain_prof, sizeof(h)));
  for (int i = 0; i < 16; ++i) out16[i] = h[i];
  if (reset) {
    unsigned long long z[16] = {};
    DSW_CUDA_TRY(cudaMemcpyToSymbol(dsw::g_chain_prof, z, sizeof(z)));
  }
  return DSW_OK;
}
