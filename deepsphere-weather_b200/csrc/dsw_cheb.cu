// Chebyshev graph convolution: forward, input gradient, weight gradient.
// Replaces conv_cheb + the bias add of ConvCheb.forward (reference modules/layers.py:113-180,
// 365-376) and what autograd derives from them.
//
// The polynomial  y = sum_k T_k(L) x W_k  can be evaluated with the sparse hops on either side of the
// dense channel mix, and the two orders cost very differently when Fin != Fout (a hop moves one
// [B,V,C] plane per channel count C):
//
//   forward   TERMS     T_k = T_k(L) x  (K-1 hops on Fin channels), then one mix (K*Fin -> Fout)   [layers.py order]
//             CLENSHAW  G_k = x W_k     (one mix Fin -> K*Fout), then the Clenshaw recurrence
//                       b_k = G_k + 2 L b_{k+1} - b_{k+2},  y = G_0 + L b_1 - b_2  (K-1 hops on Fout channels)
//   backward  CLENSHAW  G_k = dy W_k^T, Clenshaw with L^T on Fin channels -> dx; dW from the terms of x
//             TERMS     U_k = T_k(L^T) dy (K-1 hops on Fout channels); dx = sum_k U_k W_k^T;
//                       dW[f,k,o] = sum_n x[n,f] U_k[n,o]   (the adjoint form: no terms of x needed)
//
// Both orders are the same polynomial (T_k(L)^T = T_k(L^T)); the entry points pick the cheaper one
// from the channel counts (DSW_OPT_FWD_ALGO / DSW_OPT_BWD_ALGO override).  U-Net decoder layers
// (Fin = 2 Fout) run their hops on half the channels in both directions; the 64 -> 2 output layer
// on 2 instead of 64.
//
// Every entry point can also walk the batch in *sample chunks* (DSW_OPT_L2_CHUNK_BYTES) sized so that
// a chunk's intermediate planes stay resident in the 126 MB L2 between the kernels that produce and
// consume them.  Samples are independent through the whole path (layers.py:158-173), so chunking
// changes no arithmetic.  Off by default (measured slower: the per-chunk launches are too small).
#include <algorithm>

#include "dsw_internal.cuh"

namespace dsw {
// tcgen05 channel mix (dsw_mix_tc.cu).  Returns DSW_ERR_UNSUPPORTED when the shape is not taken.
size_t mix_tc_workspace_bytes(int32_t P, int32_t Ka, int32_t Nc);
int launch_mix_tc_ws(const MixArgs& a, void* prep, size_t prep_bytes, bool do_prep, cudaStream_t st);
// tcgen05 weight gradient (dsw_wgrad_tc.cu) and the fixed-order reduction of partials
int wgrad_tc_nsplit(int64_t N, int32_t Ka, int32_t Kb, int32_t Fin, int32_t Fout);
int launch_wgrad_tc(const WgradArgs& a, size_t partial_bytes, cudaStream_t st);
int launch_wgrad_reduce(const float* partial, int32_t nsplit, int32_t K, int32_t Fin, int32_t Fout, float* dW,
                        float* dbias, cudaStream_t st);

enum { ALGO_TERMS = 1, ALGO_CLENSHAW = 2 };

static int launch_mix(const MixArgs& a, void* prep, size_t prep_bytes, bool do_prep, cudaStream_t st) {
  if (g_mix_mode.load(std::memory_order_relaxed) == 1) {
    const int rc = launch_mix_tc_ws(a, prep, prep_bytes, do_prep, st);
    if (rc != DSW_ERR_UNSUPPORTED) return rc;
  }
  return launch_mix_simt(a, st);
}

static int check_common(const dsw_plan* lap, int32_t B, int32_t Fin, int32_t Fout, int32_t K) {
  if (!lap || B <= 0 || Fin <= 0 || Fout <= 0 || K < 1) return DSW_ERR_BAD_ARGUMENT;
  if (K > DSW_MAX_K) return DSW_ERR_UNSUPPORTED;
  if (lap->fwd.n_rows != lap->fwd.n_cols) return DSW_ERR_SHAPE;
  return DSW_OK;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Cost model (planes of traffic per channel, K = 4): a plain hop chain costs ~8 C, a Clenshaw chain
// ~11 C, each mix / weight gradient its operand channels.  TERMS forward 12 Fin + Fout vs CLENSHAW
// Fin + 15 Fout  =>  CLENSHAW when 14 Fout < 11 Fin; symmetrically for the backward.
static int fwd_algo(int32_t Fin, int32_t Fout, int32_t K) {
  const int64_t opt = g_options[DSW_OPT_FWD_ALGO].load(std::memory_order_relaxed);
  if (K == 1) return ALGO_TERMS;
  if (opt == ALGO_TERMS || opt == ALGO_CLENSHAW) return (int)opt;
  return (14 * (int64_t)Fout < 11 * (int64_t)Fin) ? ALGO_CLENSHAW : ALGO_TERMS;
}
static int bwd_algo(int32_t Fin, int32_t Fout, int32_t K) {
  const int64_t opt = g_options[DSW_OPT_BWD_ALGO].load(std::memory_order_relaxed);
  if (K == 1) return ALGO_CLENSHAW;
  if (opt == ALGO_TERMS || opt == ALGO_CLENSHAW) return (int)opt;
  return (14 * (int64_t)Fin < 11 * (int64_t)Fout) ? ALGO_CLENSHAW : ALGO_TERMS;
}

// Samples per chunk: the largest divisor of B whose live working set (bytes_per_sample each) fits the
// L2 budget.  DSW_OPT_L2_CHUNK_BYTES: 0 or 1 = chunking off (default), else the budget in bytes.
static int32_t chunk_samples(int32_t B, int64_t bytes_per_sample) {
  const int64_t opt = g_options[DSW_OPT_L2_CHUNK_BYTES].load(std::memory_order_relaxed);
  if (opt <= 1) return B;
  int64_t c = std::max<int64_t>(1, opt / std::max<int64_t>(bytes_per_sample, 1));
  if (c >= B) return B;
  int32_t best = 1;
  for (int32_t d = 1; d <= B && d <= c; ++d)
    if (B % d == 0) best = d;
  return best;
}
static int32_t conv_chunk(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K) {
  return chunk_samples(B, ((int64_t)K * std::min(Fin, Fout) + std::max(Fin, Fout)) * V * (int64_t)sizeof(float));
}

// T_1 .. T_{K-1} of `Bc` samples under operator A into `terms` ([K-1][plane]); T_0 is x.
static int run_terms(const dsw_csr& A, const dsw_rb& rb, const float* x, int64_t x_sB, int64_t x_sV, float* terms,
                     int64_t plane, int32_t Bc, int32_t F, int32_t K, cudaStream_t st) {
  const int64_t V = A.n_rows;
  ChainHop hops[DSW_MAX_K];
  for (int k = 1; k < K; ++k) {
    ChainHop& a = hops[k - 1];
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
  }
  return launch_hop_chain(A, rb, hops, K - 1, Bc, F, st);
}

// Clenshaw recurrence under operator A, in place on the K planes G ([K][plane], F channels):
//   b_{K-1} = G_{K-1};  b_k = G_k + 2 A b_{k+1} - b_{k+2}  (k = K-2 .. 1);  out = G_0 + A b_1 - b_2
static int run_clenshaw(const dsw_csr& A, const dsw_rb& rb, float* G, int64_t plane, float* out, int32_t Bc, int32_t F,
                        int32_t K, cudaStream_t st, int32_t act = 0, const float* mask = nullptr, int64_t m_sB = 0,
                        int64_t m_sV = 0) {
  const int64_t V = A.n_rows, sB = V * F, sV = F;
  ChainHop hops[DSW_MAX_K];
  int n = 0;
  for (int k = K - 2; k >= 0; --k) {
    ChainHop& a = hops[n++];
    a.X = G + (k + 1) * plane, a.x_sB = sB, a.x_sV = sV;
    if (k + 2 <= K - 1) a.Z = G + (k + 2) * plane, a.z_sB = sB, a.z_sV = sV, a.beta = -1.f;
    a.G = G + k * plane, a.g_sB = sB, a.g_sV = sV;
    a.alpha = (k == 0) ? 1.f : 2.f;
    a.O = (k == 0) ? out : G + k * plane, a.o_sB = sB, a.o_sV = sV;
    a.act = (k == 0) ? act : 0;  // the activation follows the last hop
    if (k == 0 && mask) a.M = mask, a.m_sB = m_sB, a.m_sV = m_sV;  // so does the ReLU mask of an input gradient
  }
  return launch_hop_chain(A, rb, hops, n, Bc, F, st);
}

// ---- weight-gradient geometry shared by the workspace queries and the calls ----
struct WgradPlan {
  int32_t nsplit;
  bool tc;
  size_t part_bytes;  // nsplit * (K*Fin + 1) * Fout floats
};
static WgradPlan wgrad_plan(int64_t Nc, int32_t Ka, int32_t Kb, int32_t Fin, int32_t Fout) {
  WgradPlan p;
  p.tc = g_mix_mode.load(std::memory_order_relaxed) == 1 && Nc < ((int64_t)1 << 31);
  p.nsplit = p.tc ? wgrad_tc_nsplit(Nc, Ka, Kb, Fin, Fout) : wgrad_pick_nsplit(Nc, Ka, Kb, Fin, Fout);
  p.part_bytes = (size_t)p.nsplit * ((size_t)Ka * Kb * Fin + 1) * Fout * sizeof(float);
  return p;
}
static int run_wgrad(const WgradArgs& w, const WgradPlan& p, cudaStream_t st) {
  if (p.tc) return launch_wgrad_tc(w, p.part_bytes, st);
  return launch_wgrad_simt(w, st);
}
}  // namespace dsw

using namespace dsw;

extern "C" {

int dsw_cheb_fwd_algo(int32_t Fin, int32_t Fout, int32_t K) { return fwd_algo(Fin, Fout, K); }
int dsw_cheb_bwd_algo(int32_t Fin, int32_t Fout, int32_t K) { return bwd_algo(Fin, Fout, K); }

int dsw_cheb_terms(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, float* terms, int32_t B,
                   int32_t F, int32_t K, void* stream) {
  if (!lap || !x || B <= 0 || F <= 0 || K < 1) return DSW_ERR_BAD_ARGUMENT;
  if (K > DSW_MAX_K) return DSW_ERR_UNSUPPORTED;
  if (lap->fwd.n_rows != lap->fwd.n_cols) return DSW_ERR_SHAPE;
  if (K > 1 && !terms) return DSW_ERR_BAD_ARGUMENT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t V = lap->fwd.n_rows, plane = (int64_t)B * V * F;
  // live set of one hop: gather source, the k-2 term, the output
  const int32_t Bc = chunk_samples(B, 3 * V * F * (int64_t)sizeof(float));
  for (int32_t b0 = 0; b0 < B; b0 += Bc) {
    // the output keeps the [K-1][B][V][F] layout: chunk b0 of term k lives at terms + (k-1)*plane + b0*V*F
    DSW_TRY(run_terms(lap->fwd, lap->fwd_rb, x + b0 * x_sB, x_sB, x_sV, terms + b0 * V * F, plane, Bc, F, K, st));
  }
  return DSW_OK;
}

size_t dsw_cheb_fwd_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K) {
  if (B <= 0 || V <= 0 || Fin <= 0 || Fout <= 0 || K < 1) return 0;
  const int32_t Bc = conv_chunk(B, V, Fin, Fout, K);
  if (fwd_algo(Fin, Fout, K) == ALGO_TERMS)
    return align_up((size_t)(K - 1) * Bc * V * Fin * sizeof(float), 256) + align_up(mix_tc_workspace_bytes(K, Fin, Fout), 256) + 256;
  return align_up((size_t)K * Bc * V * Fout * sizeof(float), 256) + align_up(mix_tc_workspace_bytes(1, Fin, K * Fout), 256) + 256;
}

int dsw_cheb_fwd(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, const float* W,
                 const float* bias, float* y, int32_t B, int32_t Fin, int32_t Fout, int32_t K, int32_t act,
                 void* workspace, size_t workspace_bytes, void* stream) {
  DSW_TRY(check_common(lap, B, Fin, Fout, K));
  if (!x || !W || !y || act < 0 || act > 1) return DSW_ERR_BAD_ARGUMENT;
  const int32_t V = lap->fwd.n_rows;
  if (!workspace || workspace_bytes < dsw_cheb_fwd_workspace_bytes(B, V, Fin, Fout, K)) return DSW_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int32_t Bc = conv_chunk(B, V, Fin, Fout, K);
  const int algo = fwd_algo(Fin, Fout, K);

  if (algo == ALGO_TERMS) {
    float* terms = static_cast<float*>(workspace);
    const size_t terms_bytes = align_up((size_t)(K - 1) * Bc * V * Fin * sizeof(float), 256);
    void* prep = static_cast<char*>(workspace) + terms_bytes;
    const size_t prep_bytes = workspace_bytes - terms_bytes;
    const int64_t plane = (int64_t)Bc * V * Fin;
    for (int32_t b0 = 0; b0 < B; b0 += Bc) {
      const float* xc = x + b0 * x_sB;
      DSW_TRY(run_terms(lap->fwd, lap->fwd_rb, xc, x_sB, x_sV, terms, plane, Bc, Fin, K, st));
      MixArgs m;
      m.P = K, m.Ka = Fin, m.rows_per_batch = V, m.N = (int64_t)Bc * V;
      m.A[0] = xc, m.a_sB[0] = x_sB, m.a_sV[0] = x_sV;
      for (int k = 1; k < K; ++k) m.A[k] = terms + (k - 1) * plane, m.a_sB[k] = (int64_t)V * Fin, m.a_sV[k] = Fin;
      // W[f][k][o]: plane k -> +k*Fout, reduction index f -> stride K*Fout, column o -> stride 1
      m.Bm = W, m.sBp = Fout, m.sBk = (int64_t)K * Fout, m.sBc0 = 1, m.sBc1 = 0;
      m.bias = bias, m.bias_n = Fout, m.C = y + (int64_t)b0 * V * Fout, m.sCp = 0, m.ldc = Fout, m.Cw = Fout, m.Nc = Fout;
      m.act = act;
      DSW_TRY(launch_mix(m, prep, prep_bytes, b0 == 0, st));
    }
    return DSW_OK;
  }

  // CLENSHAW: G_k = x W_k (+ bias on plane 0), then the recurrence with L on Fout channels
  float* G = static_cast<float*>(workspace);
  const size_t g_bytes = align_up((size_t)K * Bc * V * Fout * sizeof(float), 256);
  void* prep = static_cast<char*>(workspace) + g_bytes;
  const size_t prep_bytes = workspace_bytes - g_bytes;
  const int64_t plane = (int64_t)Bc * V * Fout;
  for (int32_t b0 = 0; b0 < B; b0 += Bc) {
    MixArgs m;
    m.P = 1, m.Ka = Fin, m.rows_per_batch = V, m.N = (int64_t)Bc * V;
    m.A[0] = x + b0 * x_sB, m.a_sB[0] = x_sB, m.a_sV[0] = x_sV;
    // output column c = k*Fout + o  <-  W[f][k][o] = W + f*K*Fout + c
    m.Bm = W, m.sBp = 0, m.sBk = (int64_t)K * Fout, m.sBc0 = 1, m.sBc1 = Fout;
    m.bias = bias, m.bias_n = Fout;  // only plane 0 (c < Fout) carries the bias
    m.C = G, m.sCp = plane, m.ldc = Fout, m.Cw = Fout, m.Nc = K * Fout, m.act = 0;
    DSW_TRY(launch_mix(m, prep, prep_bytes, b0 == 0, st));
    DSW_TRY(run_clenshaw(lap->fwd, lap->fwd_rb, G, plane, y + (int64_t)b0 * V * Fout, Bc, Fout, K, st, act));
  }
  return DSW_OK;
}

// ---------------------------------------------------------------------------------------------------
// Backward.  dsw_cheb_bwd computes dx and / or (dW, dbias) in one call so that the TERMS order shares
// its K-1 hops on dy between the two; dsw_cheb_bwd_data / dsw_cheb_bwd_weight are thin wrappers.
// ---------------------------------------------------------------------------------------------------
struct BwdLayout {
  int algo;
  int32_t Bc, nchunks;
  size_t planes_bytes;   // CLENSHAW: K planes of Fin (G) ; TERMS: K-1 planes of Fout (U)
  size_t xterms_bytes;   // CLENSHAW without saved terms: K-1 planes of Fin
  size_t prep_bytes;
  WgradPlan wg;
};
static BwdLayout bwd_layout(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K, bool need_xterms) {
  BwdLayout L;
  L.algo = bwd_algo(Fin, Fout, K);
  L.Bc = conv_chunk(B, V, Fin, Fout, K);
  L.nchunks = B / L.Bc;
  const int64_t Nc = (int64_t)L.Bc * V;
  if (L.algo == ALGO_CLENSHAW) {
    L.planes_bytes = align_up((size_t)(K > 1 ? K : 0) * Nc * Fin * sizeof(float), 256);
    L.xterms_bytes = need_xterms ? align_up((size_t)(K - 1) * Nc * Fin * sizeof(float), 256) : 0;
    L.prep_bytes = align_up(mix_tc_workspace_bytes(1, Fout, K * Fin), 256);
    L.wg = wgrad_plan(Nc, K, 1, Fin, Fout);
  } else {
    L.planes_bytes = align_up((size_t)(K - 1) * Nc * Fout * sizeof(float), 256);
    L.xterms_bytes = 0;
    L.prep_bytes = align_up(mix_tc_workspace_bytes(K, Fout, Fin), 256);
    L.wg = wgrad_plan(Nc, 1, K, Fin, Fout);
  }
  return L;
}

size_t dsw_cheb_bwd_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K, int32_t have_saved_terms) {
  if (B <= 0 || V <= 0 || Fin <= 0 || Fout <= 0 || K < 1) return 0;
  const BwdLayout L = bwd_layout(B, V, Fin, Fout, K, have_saved_terms == 0);
  return L.planes_bytes + L.xterms_bytes + L.prep_bytes + align_up(L.wg.part_bytes * L.nchunks, 256) + 256;
}

int dsw_cheb_bwd(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, const float* dy, const float* W,
                 const float* saved_terms, float* dx, float* dW, float* dbias, int32_t B, int32_t Fin, int32_t Fout,
                 int32_t K, void* workspace, size_t workspace_bytes, void* stream) {
  return dsw_cheb_bwd_ex(lap, x, x_sB, x_sV, dy, W, saved_terms, dx, dW, dbias, B, Fin, Fout, K, 0, workspace, workspace_bytes,
                         stream);
}

int dsw_cheb_bwd_ex(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, const float* dy, const float* W,
                    const float* saved_terms, float* dx, float* dW, float* dbias, int32_t B, int32_t Fin, int32_t Fout,
                    int32_t K, int32_t flags, void* workspace, size_t workspace_bytes, void* stream) {
  DSW_TRY(check_common(lap, B, Fin, Fout, K));
  if (!dy || (!dx && !dW) || (dx && !W) || (dW && !x) || (flags & ~DSW_BWD_MASK_DX_BY_X)) return DSW_ERR_BAD_ARGUMENT;
  // x = relu(.) upstream: its gradient is dx * [x > 0]; applied where the last kernel writes dx
  const bool mask_dx = (flags & DSW_BWD_MASK_DX_BY_X) && dx;
  if (mask_dx && !x) return DSW_ERR_BAD_ARGUMENT;
  const int32_t V = lap->fwd.n_rows;
  // the channel mix indexes its epilogue operand by the flat row n = b * V + v
  if (mask_dx && B > 1 && x_sB != (int64_t)V * x_sV) return DSW_ERR_UNSUPPORTED;
  const bool need_xterms = dW != nullptr && saved_terms == nullptr;
  const BwdLayout L = bwd_layout(B, V, Fin, Fout, K, need_xterms);
  if (!workspace || workspace_bytes < dsw_cheb_bwd_workspace_bytes(B, V, Fin, Fout, K, need_xterms ? 0 : 1))
    return DSW_ERR_WORKSPACE;
  if (saved_terms && (L.nchunks != 1 || L.algo != ALGO_CLENSHAW)) return DSW_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  float* planes = reinterpret_cast<float*>(ws);
  float* xterms = reinterpret_cast<float*>(ws + L.planes_bytes);
  void* prep = ws + L.planes_bytes + L.xterms_bytes;
  float* partial = reinterpret_cast<float*>(ws + L.planes_bytes + L.xterms_bytes + L.prep_bytes);
  const int64_t part_floats = (int64_t)(L.wg.part_bytes / sizeof(float));
  const int32_t Bc = L.Bc;

  for (int32_t c = 0; c < L.nchunks; ++c) {
    const int32_t b0 = c * Bc;
    const float* dyc = dy + (int64_t)b0 * V * Fout;
    const float* xc = x ? x + b0 * x_sB : nullptr;
    float* dxc = dx ? dx + (int64_t)b0 * V * Fin : nullptr;
    WgradArgs w;
    w.K = K, w.Fin = Fin, w.Fout = Fout, w.rows_per_batch = V, w.N = (int64_t)Bc * V;
    w.dW = dW, w.dbias = dbias, w.partial = partial + c * part_floats, w.nsplit = L.wg.nsplit;

    if (L.algo == ALGO_CLENSHAW) {
      if (dxc) {
        const int64_t plane = (int64_t)Bc * V * Fin;
        float* G = (K > 1) ? planes : dxc;
        // G_k[n][f] = sum_o dy[n][o] W[f][k][o]   for all k at once: output column c = k*Fin + f
        MixArgs m;
        m.P = 1, m.Ka = Fout, m.rows_per_batch = V, m.N = (int64_t)Bc * V;
        m.A[0] = dyc, m.a_sB[0] = (int64_t)V * Fout, m.a_sV[0] = Fout;
        m.Bm = W, m.sBp = 0, m.sBk = 1, m.sBc0 = (int64_t)K * Fout, m.sBc1 = Fout;
        m.bias = nullptr, m.C = G, m.sCp = plane, m.ldc = Fin, m.Cw = Fin, m.Nc = K * Fin, m.act = 0;
        if (mask_dx && K == 1) m.R = xc, m.ldr = x_sV, m.r_mode = 1;
        DSW_TRY(launch_mix(m, prep, L.prep_bytes, c == 0, st));
        if (K > 1)
          DSW_TRY(run_clenshaw(lap->tr, lap->tr_rb, G, plane, dxc, Bc, Fin, K, st, 0, mask_dx ? xc : nullptr, x_sB, x_sV));
      }
      if (dW) {
        const int64_t plane = (int64_t)Bc * V * Fin;
        const float* terms = saved_terms;
        if (!terms && K > 1) {
          DSW_TRY(run_terms(lap->fwd, lap->fwd_rb, xc, x_sB, x_sV, xterms, plane, Bc, Fin, K, st));
          terms = xterms;
        }
        w.Ka = K, w.Kb = 1;
        w.T[0] = xc, w.t_sB[0] = x_sB, w.t_sV[0] = x_sV;
        for (int k = 1; k < K; ++k) w.T[k] = terms + (k - 1) * plane, w.t_sB[k] = (int64_t)V * Fin, w.t_sV[k] = Fin;
        w.Y[0] = dyc;
        DSW_TRY(run_wgrad(w, L.wg, st));
      }
    } else {
      // U_k = T_k(L^T) dy on Fout channels, shared by dx and dW
      const int64_t plane = (int64_t)Bc * V * Fout;
      float* U = planes;
      DSW_TRY(run_terms(lap->tr, lap->tr_rb, dyc, (int64_t)V * Fout, Fout, U, plane, Bc, Fout, K, st));
      if (dxc) {
        // dx[n][f] = sum_k sum_o U_k[n][o] W[f][k][o]
        MixArgs m;
        m.P = K, m.Ka = Fout, m.rows_per_batch = V, m.N = (int64_t)Bc * V;
        m.A[0] = dyc, m.a_sB[0] = (int64_t)V * Fout, m.a_sV[0] = Fout;
        for (int k = 1; k < K; ++k) m.A[k] = U + (k - 1) * plane, m.a_sB[k] = (int64_t)V * Fout, m.a_sV[k] = Fout;
        m.Bm = W, m.sBp = Fout, m.sBk = 1, m.sBc0 = (int64_t)K * Fout, m.sBc1 = 0;
        m.bias = nullptr, m.C = dxc, m.sCp = 0, m.ldc = Fin, m.Cw = Fin, m.Nc = Fin, m.act = 0;
        if (mask_dx) m.R = xc, m.ldr = x_sV, m.r_mode = 1;
        DSW_TRY(launch_mix(m, prep, L.prep_bytes, c == 0, st));
      }
      if (dW) {
        w.Ka = 1, w.Kb = K;
        w.T[0] = xc, w.t_sB[0] = x_sB, w.t_sV[0] = x_sV;
        w.Y[0] = dyc;
        for (int k = 1; k < K; ++k) w.Y[k] = U + (k - 1) * plane;
        DSW_TRY(run_wgrad(w, L.wg, st));
      }
    }
  }
  if (dW) return launch_wgrad_reduce(partial, L.nchunks * L.wg.nsplit, K, Fin, Fout, dW, dbias, st);
  return DSW_OK;
}

size_t dsw_cheb_bwd_data_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K) {
  return dsw_cheb_bwd_workspace_bytes(B, V, Fin, Fout, K, 1);
}
int dsw_cheb_bwd_data(const dsw_plan* lap, const float* dy, const float* W, float* dx, int32_t B, int32_t Fin,
                      int32_t Fout, int32_t K, void* workspace, size_t workspace_bytes, void* stream) {
  if (!dx) return DSW_ERR_BAD_ARGUMENT;
  return dsw_cheb_bwd(lap, nullptr, 0, 0, dy, W, nullptr, dx, nullptr, nullptr, B, Fin, Fout, K, workspace, workspace_bytes,
                      stream);
}

size_t dsw_cheb_bwd_weight_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K) {
  return dsw_cheb_bwd_workspace_bytes(B, V, Fin, Fout, K, 0);
}
int dsw_cheb_bwd_weight(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, const float* dy,
                        const float* saved_terms, float* dW, float* dbias, int32_t B, int32_t Fin, int32_t Fout,
                        int32_t K, void* workspace, size_t workspace_bytes, void* stream) {
  if (!dW) return DSW_ERR_BAD_ARGUMENT;
  // saved terms of x only serve the CLENSHAW order; the TERMS order derives dW from the terms of dy
  if (saved_terms && bwd_algo(Fin, Fout, K) != ALGO_CLENSHAW) saved_terms = nullptr;
  return dsw_cheb_bwd(lap, x, x_sB, x_sV, dy, nullptr, saved_terms, nullptr, dW, dbias, B, Fin, Fout, K, workspace,
                      workspace_bytes, stream);
}

}  // extern "C"
