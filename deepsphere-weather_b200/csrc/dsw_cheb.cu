// Chebyshev graph convolution: forward, input gradient, weight gradient.
// Replaces conv_cheb + the bias add of ConvCheb.forward (reference modules/layers.py:113-180,
// 365-376) and what autograd derives from them.
#include <algorithm>

#include "dsw_internal.cuh"

namespace dsw {
// tcgen05 channel mix (dsw_mix_tc.cu).  Returns DSW_ERR_UNSUPPORTED when the shape is not taken.
size_t mix_tc_workspace_bytes(int32_t P, int32_t Ka, int32_t Nc);
size_t wgrad_tc_partial_bytes(int64_t N, int32_t K, int32_t Fin, int32_t Fout);
int launch_wgrad_tc(const WgradArgs& a, size_t partial_bytes, cudaStream_t st);
int launch_mix_tc_ws(const MixArgs& a, void* prep, size_t prep_bytes, cudaStream_t st);

static int launch_mix(const MixArgs& a, void* prep, size_t prep_bytes, cudaStream_t st) {
  if (g_mix_mode.load(std::memory_order_relaxed) == 1) {
    const int rc = launch_mix_tc_ws(a, prep, prep_bytes, st);
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

// T_1 .. T_{K-1} into `terms` ([K-1][B][V][F]); T_0 is x.
static int run_terms(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, float* terms, int32_t B,
                     int32_t F, int32_t K, cudaStream_t st) {
  return dsw_cheb_terms(lap, x, x_sB, x_sV, terms, B, F, K, st);
}
}  // namespace dsw

using namespace dsw;

extern "C" {

size_t dsw_cheb_fwd_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K) {
  if (B <= 0 || V <= 0 || Fin <= 0 || Fout <= 0 || K < 1) return 0;
  return align_up((size_t)(K - 1) * B * V * Fin * sizeof(float), 256) + align_up(mix_tc_workspace_bytes(K, Fin, Fout), 256) + 256;
}

int dsw_cheb_fwd(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, const float* W,
                 const float* bias, float* y, int32_t B, int32_t Fin, int32_t Fout, int32_t K, int32_t act,
                 void* workspace, size_t workspace_bytes, void* stream) {
  DSW_TRY(check_common(lap, B, Fin, Fout, K));
  if (!x || !W || !y || act < 0 || act > 1) return DSW_ERR_BAD_ARGUMENT;
  const int32_t V = lap->fwd.n_rows;
  if (!workspace || workspace_bytes < dsw_cheb_fwd_workspace_bytes(B, V, Fin, Fout, K)) return DSW_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* terms = static_cast<float*>(workspace);
  const size_t terms_bytes = align_up((size_t)(K - 1) * B * V * Fin * sizeof(float), 256);
  void* prep = static_cast<char*>(workspace) + terms_bytes;
  const size_t prep_bytes = workspace_bytes - terms_bytes;
  DSW_TRY(run_terms(lap, x, x_sB, x_sV, terms, B, Fin, K, st));

  MixArgs m;
  m.P = K, m.Ka = Fin, m.rows_per_batch = V, m.N = (int64_t)B * V;
  const int64_t plane = (int64_t)B * V * Fin;
  m.A[0] = x, m.a_sB[0] = x_sB, m.a_sV[0] = x_sV;
  for (int k = 1; k < K; ++k) m.A[k] = terms + (k - 1) * plane, m.a_sB[k] = (int64_t)V * Fin, m.a_sV[k] = Fin;
  // W[f][k][o]: plane k -> +k*Fout, reduction index f -> stride K*Fout, column o -> stride 1
  m.Bm = W, m.sBp = Fout, m.sBk = (int64_t)K * Fout, m.sBc0 = 1, m.sBc1 = 0;
  m.bias = bias, m.C = y, m.sCp = 0, m.ldc = Fout, m.Cw = Fout, m.Nc = Fout, m.act = act;
  return launch_mix(m, prep, prep_bytes, st);
}

size_t dsw_cheb_bwd_data_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K) {
  if (B <= 0 || V <= 0 || Fin <= 0 || Fout <= 0 || K < 1) return 0;
  return align_up((size_t)(K > 1 ? K : 0) * B * V * Fin * sizeof(float), 256) +
         align_up(mix_tc_workspace_bytes(1, Fout, K * Fin), 256) + 256;
}

int dsw_cheb_bwd_data(const dsw_plan* lap, const float* dy, const float* W, float* dx, int32_t B, int32_t Fin,
                      int32_t Fout, int32_t K, void* workspace, size_t workspace_bytes, void* stream) {
  DSW_TRY(check_common(lap, B, Fin, Fout, K));
  if (!dy || !W || !dx) return DSW_ERR_BAD_ARGUMENT;
  const int32_t V = lap->fwd.n_rows;
  if (!workspace || workspace_bytes < dsw_cheb_bwd_data_workspace_bytes(B, V, Fin, Fout, K)) return DSW_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t plane = (int64_t)B * V * Fin;
  float* G = (K > 1) ? static_cast<float*>(workspace) : dx;
  const size_t g_bytes = align_up((size_t)(K > 1 ? K : 0) * B * V * Fin * sizeof(float), 256);
  void* prep = static_cast<char*>(workspace) + g_bytes;
  const size_t prep_bytes = workspace_bytes - g_bytes;

  // G_k[n][f] = sum_o dy[n][o] W[f][k][o]   for all k at once: output column c = k*Fin + f
  MixArgs m;
  m.P = 1, m.Ka = Fout, m.rows_per_batch = V, m.N = (int64_t)B * V;
  m.A[0] = dy, m.a_sB[0] = (int64_t)V * Fout, m.a_sV[0] = Fout;
  m.Bm = W, m.sBp = 0, m.sBk = 1, m.sBc0 = (int64_t)K * Fout, m.sBc1 = Fout;
  m.bias = nullptr, m.C = G, m.sCp = plane, m.ldc = Fin, m.Cw = Fin, m.Nc = K * Fin, m.act = 0;
  DSW_TRY(launch_mix(m, prep, prep_bytes, st));
  if (K == 1) return DSW_OK;

  // Adjoint (Clenshaw) recurrence with L^T, in place on the G planes:
  //   b_{K-1} = G_{K-1};  b_k = G_k + 2 L^T b_{k+1} - b_{k+2}  (k = K-2 .. 1);  dx = G_0 + L^T b_1 - b_2
  const int64_t sB = (int64_t)V * Fin, sV = Fin;
  for (int k = K - 2; k >= 0; --k) {
    HopArgs a;
    a.B = B, a.F = Fin;
    a.X = G + (k + 1) * plane, a.x_sB = sB, a.x_sV = sV;
    if (k + 2 <= K - 1) a.Z = G + (k + 2) * plane, a.z_sB = sB, a.z_sV = sV, a.beta = -1.f;
    a.G = G + k * plane, a.g_sB = sB, a.g_sV = sV;
    a.alpha = (k == 0) ? 1.f : 2.f;
    a.O = (k == 0) ? dx : G + k * plane, a.o_sB = sB, a.o_sV = sV;
    DSW_TRY(launch_hop(lap->tr, lap->tr_rb, a, st));
  }
  return DSW_OK;
}

size_t dsw_cheb_bwd_weight_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K) {
  if (B <= 0 || V <= 0 || Fin <= 0 || Fout <= 0 || K < 1) return 0;
  const size_t terms = align_up((size_t)(K - 1) * B * V * Fin * sizeof(float), 256);
  const int ns = wgrad_pick_nsplit((int64_t)B * V, K, Fin, Fout);
  size_t part = (size_t)ns * ((size_t)K * Fin + 1) * Fout * sizeof(float);
  part = std::max(part, wgrad_tc_partial_bytes((int64_t)B * V, K, Fin, Fout));
  return terms + align_up(part, 256) + 256;
}

int dsw_cheb_bwd_weight(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, const float* dy,
                        float* dW, float* dbias, int32_t B, int32_t Fin, int32_t Fout, int32_t K,
                        void* workspace, size_t workspace_bytes, void* stream) {
  DSW_TRY(check_common(lap, B, Fin, Fout, K));
  if (!x || !dy || !dW) return DSW_ERR_BAD_ARGUMENT;
  const int32_t V = lap->fwd.n_rows;
  if (!workspace || workspace_bytes < dsw_cheb_bwd_weight_workspace_bytes(B, V, Fin, Fout, K))
    return DSW_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* terms = static_cast<float*>(workspace);
  const size_t terms_bytes = align_up((size_t)(K - 1) * B * V * Fin * sizeof(float), 256);
  DSW_TRY(run_terms(lap, x, x_sB, x_sV, terms, B, Fin, K, st));

  WgradArgs w;
  w.K = K, w.Fin = Fin, w.Fout = Fout, w.rows_per_batch = V, w.N = (int64_t)B * V;
  const int64_t plane = (int64_t)B * V * Fin;
  w.T[0] = x, w.t_sB[0] = x_sB, w.t_sV[0] = x_sV;
  for (int k = 1; k < K; ++k) w.T[k] = terms + (k - 1) * plane, w.t_sB[k] = (int64_t)V * Fin, w.t_sV[k] = Fin;
  w.dY = dy, w.dW = dW, w.dbias = dbias;
  w.partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + terms_bytes);
  if (g_mix_mode.load(std::memory_order_relaxed) == 1) {
    const int rc = launch_wgrad_tc(w, workspace_bytes - terms_bytes, st);
    if (rc != DSW_ERR_UNSUPPORTED) return rc;
  }
  w.nsplit = wgrad_pick_nsplit(w.N, K, Fin, Fout);
  return launch_wgrad_simt(w, st);
}

}  // extern "C"
