// Dense per-node linear map  y[b,v,:] = x[b,v,:] . Wl^T + bias  and its gradients — the skip connection
// of the reference's ResBlock (torch.nn.Linear at my_models_graph.py:196-201, applied in :214) on the
// same tcgen05 split-bf16 kernels as the channel mix (it is the K = 1 convolution without a Laplacian).
// Weight layout is torch's: Wl[Fout][Fin].
#include <algorithm>

#include "dsw_internal.cuh"

namespace dsw {
size_t mix_tc_workspace_bytes(int32_t P, int32_t Ka, int32_t Nc);
int launch_mix_tc_ws(const MixArgs& a, void* prep, size_t prep_bytes, bool do_prep, cudaStream_t st);
int wgrad_tc_nsplit(int64_t N, int32_t Ka, int32_t Kb, int32_t Fin, int32_t Fout);
int launch_wgrad_tc(const WgradArgs& a, size_t partial_bytes, cudaStream_t st);
int launch_wgrad_reduce(const float* partial, int32_t nsplit, int32_t K, int32_t Fin, int32_t Fout, float* dW,
                        float* dbias, cudaStream_t st);

static size_t lin_align(size_t x) { return (x + 255) / 256 * 256; }

static int lin_mix(const MixArgs& a, void* prep, size_t prep_bytes, cudaStream_t st) {
  if (g_mix_mode.load(std::memory_order_relaxed) == 1) {
    const int rc = launch_mix_tc_ws(a, prep, prep_bytes, true, st);
    if (rc != DSW_ERR_UNSUPPORTED) return rc;
  }
  return launch_mix_simt(a, st);
}

constexpr int CS_SPLITS = 296;  // column-sum partials (two per SM)

// partial[split][c] = sum over the split's rows of y[n][c]   (fixed order inside a thread: deterministic)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ y, int64_t N, int32_t F, int64_t rows_per_split,
                                                     float* __restrict__ partial) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + lane;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_split, r1 = min(r0 + rows_per_split, N);
  float s = 0.f;
  if (c < F)
    for (int64_t n = r0 + w; n < r1; n += 8) s += __ldg(y + n * F + c);
  red[w][lane] = s;
  __syncthreads();
  if (w == 0 && c < F) {
    float tsum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tsum += red[i][lane];
    partial[(int64_t)blockIdx.x * F + c] = tsum;
  }
}

struct LinLayout {
  size_t prep_bytes, part_bytes, cs_bytes;
  int nsplit;
  bool tc;
};
static LinLayout lin_layout(int64_t N, int32_t Fin, int32_t Fout) {
  LinLayout L;
  L.prep_bytes = lin_align(std::max(mix_tc_workspace_bytes(1, Fin, Fout), mix_tc_workspace_bytes(1, Fout, Fin)));
  L.tc = g_mix_mode.load(std::memory_order_relaxed) == 1 && N < ((int64_t)1 << 31);
  // weight gradient with the roles swapped: "Fin side" = dy (Fout channels), "Fout side" = x (Fin channels)
  L.nsplit = L.tc ? wgrad_tc_nsplit(N, 1, 1, Fout, Fin) : wgrad_pick_nsplit(N, 1, 1, Fout, Fin);
  L.part_bytes = lin_align((size_t)L.nsplit * ((size_t)Fout + 1) * Fin * sizeof(float));
  L.cs_bytes = lin_align((size_t)CS_SPLITS * Fout * sizeof(float));
  return L;
}
}  // namespace dsw

using namespace dsw;

extern "C" {

size_t dsw_linear_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout) {
  if (B <= 0 || V <= 0 || Fin <= 0 || Fout <= 0) return 0;
  const LinLayout L = lin_layout((int64_t)B * V, Fin, Fout);
  return L.prep_bytes + L.part_bytes + L.cs_bytes + 256;
}

int dsw_linear_fwd(const float* x, int64_t x_sB, int64_t x_sV, const float* Wl, const float* bias, float* y, int32_t B,
                   int32_t V, int32_t Fin, int32_t Fout, void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !Wl || !y || B <= 0 || V <= 0 || Fin <= 0 || Fout <= 0) return DSW_ERR_BAD_ARGUMENT;
  if (!workspace || workspace_bytes < dsw_linear_workspace_bytes(B, V, Fin, Fout)) return DSW_ERR_WORKSPACE;
  const LinLayout L = lin_layout((int64_t)B * V, Fin, Fout);
  MixArgs m;
  m.P = 1, m.Ka = Fin, m.rows_per_batch = V, m.N = (int64_t)B * V;
  m.A[0] = x, m.a_sB[0] = x_sB, m.a_sV[0] = x_sV;
  // y[n][o] = sum_f x[n][f] Wl[o][f]: reduction index f -> stride 1, column o -> stride Fin
  m.Bm = Wl, m.sBp = 0, m.sBk = 1, m.sBc0 = Fin, m.sBc1 = 0;
  m.bias = bias, m.bias_n = Fout, m.C = y, m.sCp = 0, m.ldc = Fout, m.Cw = Fout, m.Nc = Fout, m.act = 0;
  return lin_mix(m, workspace, L.prep_bytes, static_cast<cudaStream_t>(stream));
}

int dsw_linear_rezero_fwd(const float* x, int64_t x_sB, int64_t x_sV, const float* Wl, const float* bias, const float* conv_out,
                          const float* scale, float* y, int32_t B, int32_t V, int32_t Fin, int32_t Fout, void* workspace,
                          size_t workspace_bytes, void* stream) {
  return dsw_linear_rezero_fwd_ld(x, x_sB, x_sV, Wl, bias, conv_out, scale, y, Fout, B, V, Fin, Fout, workspace, workspace_bytes,
                                  stream);
}

int dsw_linear_rezero_fwd_ld(const float* x, int64_t x_sB, int64_t x_sV, const float* Wl, const float* bias, const float* conv_out,
                             const float* scale, float* y, int64_t y_ld, int32_t B, int32_t V, int32_t Fin, int32_t Fout,
                             void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !Wl || !y || !conv_out || !scale || B <= 0 || V <= 0 || Fin <= 0 || Fout <= 0 || y_ld < Fout) return DSW_ERR_BAD_ARGUMENT;
  if (!workspace || workspace_bytes < dsw_linear_workspace_bytes(B, V, Fin, Fout)) return DSW_ERR_WORKSPACE;
  const LinLayout L = lin_layout((int64_t)B * V, Fin, Fout);
  MixArgs m;
  m.P = 1, m.Ka = Fin, m.rows_per_batch = V, m.N = (int64_t)B * V;
  m.A[0] = x, m.a_sB[0] = x_sB, m.a_sV[0] = x_sV;
  m.Bm = Wl, m.sBp = 0, m.sBk = 1, m.sBc0 = Fin, m.sBc1 = 0;
  m.bias = bias, m.bias_n = Fout, m.C = y, m.sCp = 0, m.ldc = y_ld, m.Cw = Fout, m.Nc = Fout, m.act = 0;
  m.R = conv_out, m.ldr = Fout, m.r_scale = scale;  // y = x . Wl^T + bias + scale * conv_out
  return lin_mix(m, workspace, L.prep_bytes, static_cast<cudaStream_t>(stream));
}

int dsw_linear_bwd(const float* x, int64_t x_sB, int64_t x_sV, const float* dy, const float* Wl, float* dx, float* dW,
                   float* dbias, int32_t B, int32_t V, int32_t Fin, int32_t Fout, void* workspace, size_t workspace_bytes,
                   void* stream) {
  return dsw_linear_bwd_acc(x, x_sB, x_sV, dy, Wl, nullptr, dx, dW, dbias, B, V, Fin, Fout, workspace, workspace_bytes, stream);
}

int dsw_linear_bwd_acc(const float* x, int64_t x_sB, int64_t x_sV, const float* dy, const float* Wl, const float* dx_addend,
                       float* dx, float* dW, float* dbias, int32_t B, int32_t V, int32_t Fin, int32_t Fout, void* workspace,
                       size_t workspace_bytes, void* stream) {
  if (!dy || B <= 0 || V <= 0 || Fin <= 0 || Fout <= 0 || (dx_addend && !dx)) return DSW_ERR_BAD_ARGUMENT;
  if ((dx && !Wl) || (dW && !x)) return DSW_ERR_BAD_ARGUMENT;
  if (!workspace || workspace_bytes < dsw_linear_workspace_bytes(B, V, Fin, Fout)) return DSW_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t N = (int64_t)B * V;
  const LinLayout L = lin_layout(N, Fin, Fout);
  char* ws = static_cast<char*>(workspace);
  if (dx) {
    MixArgs m;
    m.P = 1, m.Ka = Fout, m.rows_per_batch = V, m.N = N;
    m.A[0] = dy, m.a_sB[0] = (int64_t)V * Fout, m.a_sV[0] = Fout;
    // dx[n][f] = sum_o dy[n][o] Wl[o][f]: reduction index o -> stride Fin, column f -> stride 1
    m.Bm = Wl, m.sBp = 0, m.sBk = Fin, m.sBc0 = 1, m.sBc1 = 0;
    m.bias = nullptr, m.C = dx, m.sCp = 0, m.ldc = Fin, m.Cw = Fin, m.Nc = Fin, m.act = 0;
    m.R = dx_addend, m.ldr = Fin;  // dx = dy . Wl + dx_addend (the convolution branch's input gradient)
    DSW_TRY(lin_mix(m, ws, L.prep_bytes, st));
  }
  if (dW) {
    // dW[o][f] = sum_n dy[n][o] x[n][f]  ==  the K = 1 weight gradient with dy on the row side
    WgradArgs w;
    w.K = 1, w.Ka = 1, w.Kb = 1, w.Fin = Fout, w.Fout = Fin, w.rows_per_batch = V, w.N = N;
    w.T[0] = dy, w.t_sB[0] = (int64_t)V * Fout, w.t_sV[0] = Fout;
    if (x_sV != Fin || (B > 1 && x_sB != (int64_t)V * Fin)) return DSW_ERR_UNSUPPORTED;  // x must be contiguous here
    w.Y[0] = x;
    w.dW = dW, w.dbias = nullptr;
    w.partial = reinterpret_cast<float*>(ws + L.prep_bytes);
    w.nsplit = L.nsplit;
    if (L.tc)
      DSW_TRY(launch_wgrad_tc(w, L.part_bytes, st));
    else
      DSW_TRY(launch_wgrad_simt(w, st));
    DSW_TRY(launch_wgrad_reduce(w.partial, L.nsplit, 1, Fout, Fin, dW, nullptr, st));
  }
  if (dbias) {
    float* part = reinterpret_cast<float*>(ws + L.prep_bytes + L.part_bytes);
    const int64_t rps = (N + CS_SPLITS - 1) / CS_SPLITS;
    const int splits = (int)((N + rps - 1) / rps);
    dim3 grid(splits, (Fout + 31) / 32);
    colsum_kernel<<<grid, 256, 0, st>>>(dy, N, Fout, rps, part);
    DSW_TRY(check_launch());
    // partial layout [splits][0 * Fout + 1][Fout]: the reduction kernel's bias row
    DSW_TRY(launch_wgrad_reduce(part, splits, 0, 0, Fout, nullptr, dbias, st));
  }
  return DSW_OK;
}

}  // extern "C"
