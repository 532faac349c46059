// Autoregressive input stacking: the step of the training loop just before model(X).
//
// The reference hands the loop to xforecasting.AutoregressiveTraining (scripts_training/train_predict_state.py:392-436):
// for every autoregressive iteration it builds  X = cat(dynamic history, boundary conditions, static)  along the feature
// axis, where the most recent slots of the dynamic history are the model's own previous predictions
// (ar_settings: input_k = [-3, -2, -1], output_k = [0], stack_most_recent_prediction; modules/utils_config.py:82-86), calls
// Y = model(X) and accumulates the weighted loss.  With torch that is a shift of the history (cat), an expand of the
// static fields and a three-way cat per iteration.  Here the history is never materialised: every input time slot is a
// POINTER (to an observed state or to an earlier prediction), and one kernel writes X[b][t][v][:] from the slots.
#include "dsw_internal.cuh"

namespace dsw {
namespace {

struct ArArgs {
  const float* dyn[DSW_AR_MAX_SLOTS];
  const float* bc[DSW_AR_MAX_SLOTS];
  float* ddyn[DSW_AR_MAX_SLOTS];
  int64_t dyn_sB[DSW_AR_MAX_SLOTS], dyn_sV[DSW_AR_MAX_SLOTS];
  int64_t bc_sB[DSW_AR_MAX_SLOTS], bc_sV[DSW_AR_MAX_SLOTS];
  const float* stat;  // [V][Fs] or null
  int32_t B, T, V, Fd, Fb, Fs;
};

// X[b][t][v][f] = f < Fd ? dyn[t][b][v][f] : f < Fd + Fb ? bc[t][b][v][f - Fd] : static[v][f - Fd - Fb]
__global__ void __launch_bounds__(256) ar_stack_kernel(const ArArgs a, float* __restrict__ X) {
  pdl_trigger();
  pdl_wait();
  const int F = a.Fd + a.Fb + a.Fs;
  const int64_t n = (int64_t)a.B * a.T * a.V * F;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int f = (int)(i % F);
    const int64_t row = i / F;
    const int v = (int)(row % a.V);
    const int64_t bt = row / a.V;
    const int t = (int)(bt % a.T), b = (int)(bt / a.T);
    float x;
    if (f < a.Fd)
      x = a.dyn[t][b * a.dyn_sB[t] + v * a.dyn_sV[t] + f];
    else if (f < a.Fd + a.Fb)
      x = a.bc[t][b * a.bc_sB[t] + v * a.bc_sV[t] + (f - a.Fd)];
    else
      x = __ldg(a.stat + (int64_t)v * a.Fs + (f - a.Fd - a.Fb));
    X[i] = x;
  }
}

// ddyn[t][b][v][f] = dX[b][t][v][f]  (f < Fd) for every slot whose source needs a gradient (dense [B][V][Fd] outputs)
__global__ void __launch_bounds__(256) ar_unstack_kernel(const ArArgs a, const float* __restrict__ dX) {
  pdl_trigger();
  pdl_wait();
  const int F = a.Fd + a.Fb + a.Fs;
  const int64_t n = (int64_t)a.B * a.T * a.V * a.Fd;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int f = (int)(i % a.Fd);
    const int64_t row = i / a.Fd;
    const int v = (int)(row % a.V);
    const int64_t bt = row / a.V;
    const int t = (int)(bt % a.T), b = (int)(bt / a.T);
    if (a.ddyn[t] != nullptr) a.ddyn[t][((int64_t)b * a.V + v) * a.Fd + f] = dX[((int64_t)(b * a.T + t) * a.V + v) * F + f];
  }
}

int fill_args(ArArgs* a, const dsw_ar_slots* s, int32_t B, int32_t T, int32_t V, int32_t Fd, int32_t Fb, int32_t Fs) {
  if (!s || B <= 0 || T <= 0 || V <= 0 || Fd <= 0 || Fb < 0 || Fs < 0) return DSW_ERR_BAD_ARGUMENT;
  if (T > DSW_AR_MAX_SLOTS) return DSW_ERR_UNSUPPORTED;
  *a = ArArgs{};
  for (int t = 0; t < T; ++t) {
    if (!s->dyn[t] || (Fb > 0 && !s->bc[t])) return DSW_ERR_BAD_ARGUMENT;
    a->dyn[t] = s->dyn[t], a->dyn_sB[t] = s->dyn_sB[t], a->dyn_sV[t] = s->dyn_sV[t];
    a->bc[t] = s->bc[t], a->bc_sB[t] = s->bc_sB[t], a->bc_sV[t] = s->bc_sV[t];
    a->ddyn[t] = s->ddyn[t];
  }
  if (Fs > 0 && !s->stat) return DSW_ERR_BAD_ARGUMENT;
  a->stat = s->stat;
  a->B = B, a->T = T, a->V = V, a->Fd = Fd, a->Fb = Fb, a->Fs = Fs;
  return DSW_OK;
}

}  // namespace
}  // namespace dsw

using namespace dsw;

extern "C" {

int dsw_ar_stack_fwd(const dsw_ar_slots* slots, float* X, int32_t B, int32_t T, int32_t V, int32_t Fd, int32_t Fb, int32_t Fs, void* stream) {
  ArArgs a;
  DSW_TRY(fill_args(&a, slots, B, T, V, Fd, Fb, Fs));
  if (!X) return DSW_ERR_BAD_ARGUMENT;
  const int64_t n = (int64_t)B * T * V * (Fd + Fb + Fs);
  const int blocks = (int)std::min<int64_t>(ceil_div64(n, 256), 148 * 8);
  DSW_CUDA_TRY(launch_pdl(ar_stack_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), pdl_enabled(), a, X));
  return check_launch();
}

int dsw_ar_stack_bwd(const dsw_ar_slots* slots, const float* dX, int32_t B, int32_t T, int32_t V, int32_t Fd, int32_t Fb, int32_t Fs,
                     void* stream) {
  if (!slots || !dX || B <= 0 || T <= 0 || V <= 0 || Fd <= 0 || Fb < 0 || Fs < 0) return DSW_ERR_BAD_ARGUMENT;
  if (T > DSW_AR_MAX_SLOTS) return DSW_ERR_UNSUPPORTED;
  ArArgs a{};
  bool any = false;
  for (int t = 0; t < T; ++t) a.ddyn[t] = slots->ddyn[t], any |= slots->ddyn[t] != nullptr;
  if (!any) return DSW_OK;
  a.B = B, a.T = T, a.V = V, a.Fd = Fd, a.Fb = Fb, a.Fs = Fs;
  const int64_t n = (int64_t)B * T * V * Fd;
  const int blocks = (int)std::min<int64_t>(ceil_div64(n, 256), 148 * 8);
  DSW_CUDA_TRY(launch_pdl(ar_unstack_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), pdl_enabled(), a, dX));
  return check_launch();
}

}  // extern "C"
