// Chebyshev graph convolution: forward, input gradient, weight gradient.
// Replaces conv_cheb + the bias add of ConvCheb.forward (reference modules/layers.py:113-180,
// 365-376) and what autograd derives from them.
//
// Every entry point can walk the batch in *sample chunks* (DSW_OPT_L2_CHUNK_BYTES) sized so that the
// chunk's Chebyshev terms (and, for the input gradient, its G planes) stay resident in the 126 MB L2
// between the kernels that produce and consume them: the K-1 hops and the channel mix of one chunk
// run back to back on a workspace slice that is reused for the next chunk.  Samples are independent
// through the whole path (layers.py:158-173), so chunking changes no arithmetic.  Off by default.
#include <algorithm>

#include "dsw_internal.cuh"

namespace dsw {
// tcgen05 channel mix (dsw_mix_tc.cu).  Returns DSW_ERR_UNSUPPORTED when the shape is not taken.
size_t mix_tc_workspace_bytes(int32_t P, int32_t Ka, int32_t Nc);
int launch_mix_tc_ws(const MixArgs& a, void* prep, size_t prep_bytes, bool do_prep, cudaStream_t st);
// tcgen05 weight gradient (dsw_wgrad_tc.cu) and the fixed-order reduction of partials
int wgrad_tc_nsplit(int64_t N, int32_t K, int32_t Fin, int32_t Fout);
size_t wgrad_tc_partial_bytes(int64_t N, int32_t K, int32_t Fin, int32_t Fout);
int launch_wgrad_tc(const WgradArgs& a, size_t partial_bytes, cudaStream_t st);
int launch_wgrad_reduce(const float* partial, int32_t nsplit, int32_t K, int32_t Fin, int32_t Fout, float* dW,
                        float* dbias, cudaStream_t st);

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

// Samples per chunk: the largest divisor of B whose live working set (bytes_per_sample each) fits the
// L2 budget.  DSW_OPT_L2_CHUNK_BYTES: 0 or 1 = chunking off (default: measured slower on B200 — the
// per-chunk launches are too small to fill 148 SMs), else the budget in bytes.
int32_t chunk_samples(int32_t B, int64_t bytes_per_sample) {
  const int64_t opt = g_options[DSW_OPT_L2_CHUNK_BYTES].load(std::memory_order_relaxed);
  if (opt <= 1) return B;
  int64_t c = std::max<int64_t>(1, opt / std::max<int64_t>(bytes_per_sample, 1));
  if (c >= B) return B;
  int32_t best = 1;
  for (int32_t d = 1; d <= B && d <= c; ++d)
    if (B % d == 0) best = d;
  return best;
}

// T_1 .. T_{K-1} of samples [0, Bc) into `terms` ([K-1][plane], plane = Bc*V*F); T_0 is x.
static int run_terms(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, float* terms, int64_t plane,
                     int32_t Bc, int32_t F, int32_t K, cudaStream_t st) {
  const int64_t V = lap->fwd.n_rows;
  for (int k = 1; k < K; ++k) {
    HopArgs a;
    a.B = Bc, a.F = F;
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
}  // namespace dsw

using namespace dsw;

extern "C" {

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
    DSW_TRY(run_terms(lap, x + b0 * x_sB, x_sB, x_sV, terms + b0 * V * F, plane, Bc, F, K, st));
  }
  return DSW_OK;
}

size_t dsw_cheb_fwd_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K) {
  if (B <= 0 || V <= 0 || Fin <= 0 || Fout <= 0 || K < 1) return 0;
  const int32_t Bc = chunk_samples(B, ((int64_t)K * Fin + Fout) * V * (int64_t)sizeof(float));
  return align_up((size_t)(K - 1) * Bc * V * Fin * sizeof(float), 256) + align_up(mix_tc_workspace_bytes(K, Fin, Fout), 256) + 256;
}

int dsw_cheb_fwd(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, const float* W,
                 const float* bias, float* y, int32_t B, int32_t Fin, int32_t Fout, int32_t K, int32_t act,
                 void* workspace, size_t workspace_bytes, void* stream) {
  DSW_TRY(check_common(lap, B, Fin, Fout, K));
  if (!x || !W || !y || act < 0 || act > 1) return DSW_ERR_BAD_ARGUMENT;
  const int32_t V = lap->fwd.n_rows;
  if (!workspace || workspace_bytes < dsw_cheb_fwd_workspace_bytes(B, V, Fin, Fout, K)) return DSW_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int32_t Bc = chunk_samples(B, ((int64_t)K * Fin + Fout) * V * (int64_t)sizeof(float));
  float* terms = static_cast<float*>(workspace);
  const size_t terms_bytes = align_up((size_t)(K - 1) * Bc * V * Fin * sizeof(float), 256);
  void* prep = static_cast<char*>(workspace) + terms_bytes;
  const size_t prep_bytes = workspace_bytes - terms_bytes;
  const int64_t plane = (int64_t)Bc * V * Fin;

  for (int32_t b0 = 0; b0 < B; b0 += Bc) {
    const float* xc = x + b0 * x_sB;
    DSW_TRY(run_terms(lap, xc, x_sB, x_sV, terms, plane, Bc, Fin, K, st));
    MixArgs m;
    m.P = K, m.Ka = Fin, m.rows_per_batch = V, m.N = (int64_t)Bc * V;
    m.A[0] = xc, m.a_sB[0] = x_sB, m.a_sV[0] = x_sV;
    for (int k = 1; k < K; ++k) m.A[k] = terms + (k - 1) * plane, m.a_sB[k] = (int64_t)V * Fin, m.a_sV[k] = Fin;
    // W[f][k][o]: plane k -> +k*Fout, reduction index f -> stride K*Fout, column o -> stride 1
    m.Bm = W, m.sBp = Fout, m.sBk = (int64_t)K * Fout, m.sBc0 = 1, m.sBc1 = 0;
    m.bias = bias, m.C = y + (int64_t)b0 * V * Fout, m.sCp = 0, m.ldc = Fout, m.Cw = Fout, m.Nc = Fout, m.act = act;
    DSW_TRY(launch_mix(m, prep, prep_bytes, b0 == 0, st));
  }
  return DSW_OK;
}

size_t dsw_cheb_bwd_data_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K) {
  if (B <= 0 || V <= 0 || Fin <= 0 || Fout <= 0 || K < 1) return 0;
  const int32_t Bc = chunk_samples(B, ((int64_t)K * Fin + Fout) * V * (int64_t)sizeof(float));
  return align_up((size_t)(K > 1 ? K : 0) * Bc * V * Fin * sizeof(float), 256) +
         align_up(mix_tc_workspace_bytes(1, Fout, K * Fin), 256) + 256;
}

int dsw_cheb_bwd_data(const dsw_plan* lap, const float* dy, const float* W, float* dx, int32_t B, int32_t Fin,
                      int32_t Fout, int32_t K, void* workspace, size_t workspace_bytes, void* stream) {
  DSW_TRY(check_common(lap, B, Fin, Fout, K));
  if (!dy || !W || !dx) return DSW_ERR_BAD_ARGUMENT;
  const int32_t V = lap->fwd.n_rows;
  if (!workspace || workspace_bytes < dsw_cheb_bwd_data_workspace_bytes(B, V, Fin, Fout, K)) return DSW_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int32_t Bc = chunk_samples(B, ((int64_t)K * Fin + Fout) * V * (int64_t)sizeof(float));
  const int64_t plane = (int64_t)Bc * V * Fin;
  const size_t g_bytes = align_up((size_t)(K > 1 ? K : 0) * Bc * V * Fin * sizeof(float), 256);
  void* prep = static_cast<char*>(workspace) + g_bytes;
  const size_t prep_bytes = workspace_bytes - g_bytes;
  const int64_t sB = (int64_t)V * Fin, sV = Fin;

  for (int32_t b0 = 0; b0 < B; b0 += Bc) {
    float* dxc = dx + (int64_t)b0 * V * Fin;
    float* G = (K > 1) ? static_cast<float*>(workspace) : dxc;
    // G_k[n][f] = sum_o dy[n][o] W[f][k][o]   for all k at once: output column c = k*Fin + f
    MixArgs m;
    m.P = 1, m.Ka = Fout, m.rows_per_batch = V, m.N = (int64_t)Bc * V;
    m.A[0] = dy + (int64_t)b0 * V * Fout, m.a_sB[0] = (int64_t)V * Fout, m.a_sV[0] = Fout;
    m.Bm = W, m.sBp = 0, m.sBk = 1, m.sBc0 = (int64_t)K * Fout, m.sBc1 = Fout;
    m.bias = nullptr, m.C = G, m.sCp = plane, m.ldc = Fin, m.Cw = Fin, m.Nc = K * Fin, m.act = 0;
    DSW_TRY(launch_mix(m, prep, prep_bytes, b0 == 0, st));
    if (K == 1) continue;

    // Adjoint (Clenshaw) recurrence with L^T, in place on the G planes:
    //   b_{K-1} = G_{K-1};  b_k = G_k + 2 L^T b_{k+1} - b_{k+2}  (k = K-2 .. 1);  dx = G_0 + L^T b_1 - b_2
    for (int k = K - 2; k >= 0; --k) {
      HopArgs a;
      a.B = Bc, a.F = Fin;
      a.X = G + (k + 1) * plane, a.x_sB = sB, a.x_sV = sV;
      if (k + 2 <= K - 1) a.Z = G + (k + 2) * plane, a.z_sB = sB, a.z_sV = sV, a.beta = -1.f;
      a.G = G + k * plane, a.g_sB = sB, a.g_sV = sV;
      a.alpha = (k == 0) ? 1.f : 2.f;
      a.O = (k == 0) ? dxc : G + k * plane, a.o_sB = sB, a.o_sV = sV;
      DSW_TRY(launch_hop(lap->tr, lap->tr_rb, a, st));
    }
  }
  return DSW_OK;
}

// Weight-gradient chunk geometry shared by the workspace query and the call.
struct WgradPlan {
  int32_t Bc, nchunks, nsplit;
  bool tc;
  size_t terms_bytes, part_bytes;  // per-chunk partial region = nsplit * (K*Fin + 1) * Fout floats
};
static WgradPlan wgrad_plan(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K) {
  WgradPlan p;
  p.Bc = chunk_samples(B, ((int64_t)K * Fin + Fout) * V * (int64_t)sizeof(float));
  p.nchunks = B / p.Bc;
  const int64_t Nc = (int64_t)p.Bc * V;
  p.tc = g_mix_mode.load(std::memory_order_relaxed) == 1 && Nc < ((int64_t)1 << 31);
  p.nsplit = p.tc ? wgrad_tc_nsplit(Nc, K, Fin, Fout) : wgrad_pick_nsplit(Nc, K, Fin, Fout);
  p.terms_bytes = align_up((size_t)(K - 1) * p.Bc * V * Fin * sizeof(float), 256);
  p.part_bytes = (size_t)p.nsplit * ((size_t)K * Fin + 1) * Fout * sizeof(float);
  return p;
}

size_t dsw_cheb_bwd_weight_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K) {
  if (B <= 0 || V <= 0 || Fin <= 0 || Fout <= 0 || K < 1) return 0;
  const WgradPlan p = wgrad_plan(B, V, Fin, Fout, K);
  return p.terms_bytes + align_up(p.part_bytes * p.nchunks, 256) + 256;
}

int dsw_cheb_bwd_weight(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, const float* dy,
                        const float* saved_terms, float* dW, float* dbias, int32_t B, int32_t Fin, int32_t Fout,
                        int32_t K, void* workspace, size_t workspace_bytes, void* stream) {
  DSW_TRY(check_common(lap, B, Fin, Fout, K));
  if (!x || !dy || !dW) return DSW_ERR_BAD_ARGUMENT;
  const int32_t V = lap->fwd.n_rows;
  if (!workspace || workspace_bytes < dsw_cheb_bwd_weight_workspace_bytes(B, V, Fin, Fout, K))
    return DSW_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const WgradPlan p = wgrad_plan(B, V, Fin, Fout, K);
  if (saved_terms && p.nchunks != 1) return DSW_ERR_UNSUPPORTED;  // saved terms need the unchunked layout
  float* terms = saved_terms ? const_cast<float*>(saved_terms) : static_cast<float*>(workspace);
  float* partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + p.terms_bytes);
  const int64_t plane = (int64_t)p.Bc * V * Fin;
  const int64_t part_floats = (int64_t)(p.part_bytes / sizeof(float));

  for (int32_t c = 0; c < p.nchunks; ++c) {
    const int32_t b0 = c * p.Bc;
    const float* xc = x + b0 * x_sB;
    if (!saved_terms) DSW_TRY(run_terms(lap, xc, x_sB, x_sV, terms, plane, p.Bc, Fin, K, st));
    WgradArgs w;
    w.K = K, w.Fin = Fin, w.Fout = Fout, w.rows_per_batch = V, w.N = (int64_t)p.Bc * V;
    w.T[0] = xc, w.t_sB[0] = x_sB, w.t_sV[0] = x_sV;
    for (int k = 1; k < K; ++k) w.T[k] = terms + (k - 1) * plane, w.t_sB[k] = (int64_t)V * Fin, w.t_sV[k] = Fin;
    w.dY = dy + (int64_t)b0 * V * Fout, w.dW = dW, w.dbias = dbias;
    w.partial = partial + c * part_floats;
    w.nsplit = p.nsplit;
    if (p.tc)
      DSW_TRY(launch_wgrad_tc(w, p.part_bytes, st));
    else
      DSW_TRY(launch_wgrad_simt(w, st));
  }
  // all chunk partials are contiguous: one fixed-order reduction over nchunks * nsplit of them
  return launch_wgrad_reduce(partial, p.nchunks * p.nsplit, K, Fin, Fout, dW, dbias, st);
}

}  // extern "C"
