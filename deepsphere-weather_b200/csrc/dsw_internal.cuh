// Internal declarations shared by the libdsw.so translation units (not part of the C-ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "dsw.h"

#define DSW_TILE_BLOCKS 16
#define HOP_CNT_SLOTS 64
#define DSW_PANEL_PAD 4        // zero entry steps appended to every tile's panels (the software pipeline over-reads)
#define DSW_CHAIN_SETS 8       // rotating sets of chain-kernel sync words (claim counter, epoch, tile flags)
#define DSW_CHAIN_HDR 16       // ints in front of the flags of a set: [0] claim, [1] CTAs gone, [2] epoch
#define DSW_CHAIN_MAX_DEPS 32  // a tile may gather from at most this many tiles for the fused chain kernel

struct dsw_csr {
  int32_t n_rows = 0, n_cols = 0;
  int64_t nnz = 0;
  int32_t max_row_nnz = 0;
  int32_t* rowptr = nullptr;  // [n_rows + 1]
  int32_t* col = nullptr;     // [nnz], ascending within a row
  float* val = nullptr;       // [nnz]
};

// Row-block layout ("RB"): R consecutive rows share one list of the distinct columns they touch
// (their union), and hold a dense R x U weight panel (zero where a row lacks the column).  One
// gathered feature vector then feeds R FMAs, which cuts the L1->register traffic per useful FMA
// by  sum(nnz) / U  (2.6x on HEALPix nested k-NN-20 with R = 4).
struct dsw_rb {
  int32_t R = 0;               // rows per block (0 = layout not built)
  int32_t n_blocks = 0;
  int32_t max_union = 0;
  int64_t total_union = 0;
  int32_t tile_entries_max = 0;  // max union entries over runs of 32 consecutive row-blocks
  // Tiles of DSW_TILE_BLOCKS consecutive row-blocks: the sorted list of distinct source rows a tile
  // gathers (staged in shared memory by bulk async copies) and, per union entry, its position in
  // that list.
  int32_t n_tiles = 0;
  int32_t tile_rows_max = 0;
  int32_t* tile_ptr = nullptr;   // [n_tiles + 1] offsets into tile_row
  int32_t* tile_row = nullptr;   // source row ids, ascending inside a tile
  uint16_t* lidx = nullptr;      // [total_union] local index of ucol[e] inside its tile's list
  // TMA pieces of a tile: runs of consecutive source rows cut into power-of-two lengths (one
  // cp.async.bulk.tensor box each).  meta = (first local row << 8) | log2(length).
  int32_t tile_pieces_max = 0;
  int32_t* tpc_ptr = nullptr;    // [n_tiles + 1]
  int32_t* tpc_row = nullptr;    // first source row of the piece
  uint32_t* tpc_meta = nullptr;
  // Entry-major padded panels of a tile: entry u of row-block slot s at [tp_ptr[t] + u][s]; every
  // row-block of the tile is padded to the tile's longest union with zero weights / offset 0, so the
  // 32 slots of one entry step are contiguous (conflict-free broadcast reads, uniform trip counts).
  int32_t tile_len_max = 0;      // max entry steps of a tile
  int32_t* tp_ptr = nullptr;     // [n_tiles + 1] cumulative entry steps (each tile: its longest union + DSW_PANEL_PAD zero steps)
  float4* tp_val = nullptr;      // [total_steps][32] R = 4 weights
  uint32_t* tp_off = nullptr;    // [total_steps][32] byte offset of the source row inside the staged tile (256 B rows)
  // Dynamic item scheduling of the tile hop kernel: HOP_CNT_SLOTS rotating sets of per-tile claim
  // counters + one "CTAs gone" word (all zero between launches: the last CTA of a launch clears its
  // set) and the host-side launch counter that picks the set.
  int32_t* hop_cnt = nullptr;               // [HOP_CNT_SLOTS][n_tiles + 1]
  std::atomic<uint32_t>* hop_ring = nullptr;
  // Fused multi-hop chain kernel (dsw_chain.cu): the tiles a tile gathers from (own tile included) and the
  // rotating sets of sync words [DSW_CHAIN_SETS][DSW_CHAIN_HDR + chain_flag_cap].
  int32_t tile_deps_max = 0;     // 0 = chain kernel not available for this operator
  int32_t* tdep_ptr = nullptr;   // [n_tiles + 1]
  int32_t* tdep_idx = nullptr;
  // the same per-tile metadata at fixed strides, so that the chain kernel's issuer warp fetches all of it in one round trip
  int4* tile_meta = nullptr;     // [n_tiles][2]: {panel step offset, steps incl. pad, source rows, pieces}, {deps, entry-loop steps of compute warp 0, of warp 1, 0}
  int2* tpc_fix = nullptr;       // [n_tiles][tile_pieces_max]: {meta, first source row}; meta = 0xffffffff past the tile's pieces
  int32_t* tdep_fix = nullptr;   // [n_tiles][tile_deps_max], -1 past the tile's dependencies
  int32_t chain_flag_cap = 0;
  int32_t* chain_sync = nullptr;
  std::atomic<uint32_t>* chain_ring = nullptr;
  int32_t* perm = nullptr;     // locality permutation: original row of permuted position p ([n_blocks * R], -1 = padding) or null
  int32_t* blkptr = nullptr;   // [n_blocks + 1] offsets into ucol / uval panels
  int32_t* ucol = nullptr;     // [total_union]
  float* uval = nullptr;       // [total_union * R]  (entry u, row r) at uval[u*R + r]
};

struct dsw_plan {
  int device = 0;
  dsw_csr fwd;  // the operator
  dsw_csr tr;   // its transpose
  dsw_rb fwd_rb, tr_rb;
};

namespace dsw {

extern std::atomic<int64_t> g_launches;
extern std::atomic<int> g_mix_mode;
extern std::atomic<int64_t> g_options[DSW_OPT_COUNT];

void set_cuda_error(cudaError_t e);

inline int check_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_cuda_error(e);
    return DSW_ERR_CUDA;
  }
  return DSW_OK;
}

#define DSW_CUDA_TRY(expr)              \
  do {                                  \
    cudaError_t _e = (expr);            \
    if (_e != cudaSuccess) {            \
      ::dsw::set_cuda_error(_e);        \
      return DSW_ERR_CUDA;              \
    }                                   \
  } while (0)

#define DSW_TRY(expr)          \
  do {                         \
    int _rc = (expr);          \
    if (_rc != DSW_OK) return _rc; \
  } while (0)

// DSW_OPT_CONV_MODE -> split_pair mode (0 -> packed F2FP, 1 -> F2F per value, 2 -> integer rounding)
inline int split_mode() {
  const int64_t v = g_options[DSW_OPT_CONV_MODE].load(std::memory_order_relaxed);
  return v == 1 ? 0 : v == 2 ? 2 : 1;
}

// One-time-per-DEVICE guard for cudaFuncSetAttribute (function attributes belong to a device's context: a process-wide
// "done" flag would leave the > 48 KB shared-memory kernels un-opted-in on a second GPU).  The attribute is applied
// BEFORE the bit is published, so a second host thread either sees the bit (attribute in place) or applies it again.
struct PerDeviceOnce {
  std::atomic<uint64_t> mask{0};
  template <class Kern>
  cudaError_t max_dynamic_smem(Kern kern, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t bit = 1ull << (dev & 63);
    if (mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) mask.fetch_or(bit, std::memory_order_release);
    return e;
  }
};

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- sparse hop:  out = alpha * (A . X) + beta * Z + G  on [B][rows][F] slabs -------------------
struct HopArgs {
  const float* X = nullptr;  // gather source  [B][n_cols][F]
  int64_t x_sB = 0, x_sV = 0;
  const float* Z = nullptr;  // optional, indexed like the output rows
  int64_t z_sB = 0, z_sV = 0;
  const float* G = nullptr;  // optional, indexed like the output rows
  int64_t g_sB = 0, g_sV = 0;
  float* O = nullptr;        // [B][n_rows][F]
  int64_t o_sB = 0, o_sV = 0;
  float alpha = 1.f, beta = 0.f;
  int32_t B = 0, F = 0;
  int32_t act = 0;           // 1 = ReLU on the result (the last hop of a fused conv + activation)
  const float* M = nullptr;  // optional mask, indexed like the output rows: O = (M > 0) ? O : 0
  int64_t m_sB = 0, m_sV = 0;
};
int launch_hop(const dsw_csr& A, const dsw_rb& rb, const HopArgs& a, cudaStream_t st);

// ---- chain of hops: hop j gathers from hop j-1's output (dsw_chain.cu) ---------------------------------
// One persistent launch walks all hops in an L2-resident order when the operator's plan supports it; otherwise
// the hops are launched one by one.  Z / G of hop j may be external or the output of an earlier hop of the chain.
#define DSW_CHAIN_MAX_HOPS 7
struct ChainHop {
  const float* X = nullptr;
  const float* Z = nullptr;
  const float* G = nullptr;
  float* O = nullptr;
  int64_t x_sB = 0, x_sV = 0, z_sB = 0, z_sV = 0, g_sB = 0, g_sV = 0, o_sB = 0, o_sV = 0;
  float alpha = 1.f, beta = 0.f;
  int32_t act = 0;
  int32_t dep = 0;  // set by the launcher: 1 = wait for the previous hop's tiles
  const float* M = nullptr;  // optional mask (see HopArgs)
  int64_t m_sB = 0, m_sV = 0;
};
int launch_hop_chain(const dsw_csr& A, const dsw_rb& rb, const ChainHop* hops, int n, int32_t B, int32_t F, cudaStream_t st);

// ---- dense channel mix (CUDA-core fp32) -----------------------------------------------------------
// C[n][c] = bias[c] + sum_p sum_kk A_p[n][kk] * Bm[p*sBp + kk*sBk + (c/Cw)*sBc1 + (c%Cw)*sBc0]
// written to  C + (c/Cw)*sCp + n*ldc + (c%Cw).
struct MixArgs {
  const float* A[DSW_MAX_K];  // plane base pointers (row n at A[p] + rowoff(n))
  int32_t P = 1;              // number of A planes
  int32_t Ka = 0;             // reduction length per plane
  int32_t rows_per_batch = 0; // V: row n -> (b = n / V, v = n % V)
  int64_t a_sB[DSW_MAX_K];
  int64_t a_sV[DSW_MAX_K];
  const float* Bm = nullptr;
  int64_t sBp = 0, sBk = 0, sBc0 = 1, sBc1 = 0;
  const float* bias = nullptr;
  int32_t bias_n = 0;   // bias is added to output columns c < bias_n (c indexes bias)
  float* C = nullptr;
  int64_t sCp = 0, ldc = 0;
  int32_t Cw = 0;   // width of one output plane
  int32_t Nc = 0;   // total output columns
  int64_t N = 0;    // rows
  int32_t act = 0;
  // optional addend (single-plane outputs only, Cw == Nc):  C[n][c] += r_scale[0] * R[n * ldr + c]
  const float* R = nullptr;
  int64_t ldr = 0;
  const float* r_scale = nullptr;  // device scalar; null = 1
  int32_t r_mode = 0;              // 0: C += r_scale * R;  1: C = (R > 0) ? C : 0  (ReLU mask of a gradient by the ReLU's output)
};
int launch_mix_simt(const MixArgs& a, cudaStream_t st);

// ---- weight gradient (CUDA-core fp32): dW[f][k][o] = sum_n T_k[n][f] dY[n][o] -----------------------
struct WgradArgs {
  const float* T[DSW_MAX_K];
  int64_t t_sB[DSW_MAX_K];
  int64_t t_sV[DSW_MAX_K];
  int32_t K = 0, Fin = 0, Fout = 0;
  // Ka planes on the Fin side (T[0..Ka)) x Kb planes on the Fout side (Y[0..Kb)), one of them 1:
  //   Ka = K, Kb = 1:  dW[f][k][o] = sum_n T_k[n][f] dY[n][o]            (terms of x)
  //   Ka = 1, Kb = K:  dW[f][k][o] = sum_n x[n][f] (T_k(L^T) dY)[n][o]   (terms of dY, the adjoint form)
  int32_t Ka = 0, Kb = 1;
  int32_t rows_per_batch = 0;
  int64_t N = 0;
  const float* Y[DSW_MAX_K];  // [N][Fout] contiguous planes; Y[0] is dY itself (feeds dbias)
  float* dW = nullptr;        // [Fin][K][Fout]
  float* dbias = nullptr;     // [Fout] or null
  float* partial = nullptr;   // workspace [nsplit][K*Fin + 1][Fout]
  int32_t nsplit = 0;
};
int wgrad_pick_nsplit(int64_t N, int32_t Ka, int32_t Kb, int32_t Fin, int32_t Fout);
int launch_wgrad_simt(const WgradArgs& a, cudaStream_t st);

}  // namespace dsw

#ifdef __CUDACC__
#include <cuda_bf16.h>
namespace dsw {
// fp32 -> (bf16 hi, bf16 lo) operand split of the tcgen05 dense kernels, two values at a time, packed
// (first value in the low half).  hi = RN_bf16(v), lo = RN_bf16(v - hi); v - hi is exact in fp32.
//   mode 0: one F2F.BF16.F32 per value (the slow conversion pipe: 4 per pair)
//   mode 1: F2FP.BF16.F32.PACK_AB, one instruction per pair and image (default)
//   mode 2: integer rounding (add half an ulp, keep the upper 16 bits): ALU / FMA pipes only
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo, int mode) {
  if (mode == 1) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xFFFF0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - ah, b - bh);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  } else if (mode == 2) {
    const uint32_t ua = (__float_as_uint(a) + 0x8000u) & 0xFFFF0000u, ub = (__float_as_uint(b) + 0x8000u) & 0xFFFF0000u;
    hi = __byte_perm(ua, ub, 0x7632);
    const float al = a - __uint_as_float(ua), bl = b - __uint_as_float(ub);
    lo = __byte_perm(__float_as_uint(al) + 0x8000u, __float_as_uint(bl) + 0x8000u, 0x7632);
  } else {
    const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
    const __nv_bfloat16 la = __float2bfloat16_rn(a - __bfloat162float(ha)), lb = __float2bfloat16_rn(b - __bfloat162float(hb));
    hi = (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
    lo = (uint32_t)__bfloat16_as_ushort(la) | ((uint32_t)__bfloat16_as_ushort(lb) << 16);
  }
}
__device__ __forceinline__ void split_quad(const float4& v, uint2& hi, uint2& lo, int mode) {
  split_pair(v.x, v.y, hi.x, lo.x, mode);
  split_pair(v.z, v.w, hi.y, lo.y, mode);
}
}  // namespace dsw
#endif

#ifdef __CUDACC__
namespace dsw {
// Programmatic dependent launch (PDL).  A kernel that calls pdl_trigger() lets the NEXT kernel of the stream —
// if that one was launched with launch_pdl() — start its CTAs while this grid is still finishing; such a
// dependent kernel reads only launch-invariant data (plans, barriers, TMEM) until it calls pdl_wait(), which
// returns once every earlier kernel has completed and its writes are visible.  Both are no-ops otherwise.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class Kern, class... Args>
inline cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
inline bool pdl_enabled() { return g_options[DSW_OPT_NO_PDL].load(std::memory_order_relaxed) == 0; }
}  // namespace dsw
#endif
