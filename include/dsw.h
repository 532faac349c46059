/*
 * dsw.h — C-ABI of libdsw.so: the B200 (sm_100a) implementation of DeepSphere-Weather's
 * spherical graph-convolution hot path.
 *
 * The reference (deepsphere/deepsphere-weather) is pure Python: it has no FFI seam, the seam is
 * Python class identity in modules/layers.py.  This header is therefore the
 * interface a ctypes binding inside modules/layers.py would load (INTEGRATION.md shows that
 * binding); every entry point names the reference lines it replaces (paths relative to the
 * reference root).
 *
 * Conventions
 *  - plain C types only; no torch / C++ types cross this boundary;
 *  - every data pointer is a DEVICE pointer on the current CUDA device unless stated otherwise;
 *  - `stream` is a cudaStream_t passed as void*; every kernel is launched on it, nothing syncs
 *    except dsw_plan_create (one-time setup);
 *  - the caller owns all tensors, outputs and workspaces; the library owns only dsw_plan
 *    internals.  No allocation happens on the hot path;
 *  - node features are fp32, channel-last: x[b][v][f] at x + b*sB + v*sV + f (element strides,
 *    feature stride 1).  Outputs are written densely ([B][V][F] contiguous);
 *  - return value: 0 on success, a negative dsw_status otherwise; dsw_strerror() names it.
 *  - re-entrant: calls may come from any host thread (PyTorch runs the backward on its autograd worker).  The only
 *    process-wide mutable state are the tuning switches (dsw_set_option / dsw_set_mix_mode: relaxed atomics read at
 *    launch time) and the launch counter; a plan's device arrays are immutable after creation except its scheduling
 *    words (claim counters, chain-kernel flags), which every launch leaves zeroed / advanced for the next one.
 */
#ifndef DSW_H_
#define DSW_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSW_VERSION 100 /* major*100 + minor */

typedef enum dsw_status {
  DSW_OK = 0,
  DSW_ERR_BAD_ARGUMENT = -1,  /* null pointer, non-positive size, bad enum */
  DSW_ERR_SHAPE = -2,         /* shapes inconsistent with the plan / each other */
  DSW_ERR_WORKSPACE = -3,     /* workspace pointer null or too small */
  DSW_ERR_UNSUPPORTED = -4,   /* e.g. kernel_size > DSW_MAX_K */
  DSW_ERR_CUDA = -5,          /* a CUDA runtime call or launch failed (see dsw_last_cuda_error) */
  DSW_ERR_NO_DEVICE = -6,     /* no CUDA device / wrong architecture (needs sm_100) */
  DSW_ERR_ALIGNMENT = -7      /* pointer / stride alignment required by the fast path violated */
} dsw_status;

#define DSW_MAX_K 16

/* Opaque sparse-operator plan: CSR (int32) of the operator and of its transpose, plus the
 * row-block layouts the kernels use.  Built once per Laplacian / remap matrix. */
typedef struct dsw_plan dsw_plan;

int dsw_version(void);
const char* dsw_strerror(int status);
/* Last cudaError_t (as int) recorded by a failing call on this thread, and its string. */
int dsw_last_cuda_error(void);
const char* dsw_last_cuda_error_string(void);

/* Number of CUDA devices visible (0 without a GPU; never fails). */
int dsw_device_count(void);

/* ---------------------------------------------------------------------------------------------
 * Plans.  Replaces the per-call COO->CSR conversion hidden inside torch.sparse.mm
 * (modules/layers.py:164,167,962) and consumes exactly what prepare_torch_laplacian
 * (layers.py:82-106) / convert_to_torch_sparse (layers.py:584-594) produce: a coalesced COO
 * matrix with int64 indices and fp32 values.  coo_row / coo_col / coo_val may be host or device
 * pointers.  Duplicates are summed; entries need not be sorted.
 * ------------------------------------------------------------------------------------------- */
int dsw_plan_create(int32_t n_rows, int32_t n_cols, int64_t nnz, const int64_t* coo_row,
                    const int64_t* coo_col, const float* coo_val, void* stream, dsw_plan** out);
void dsw_plan_destroy(dsw_plan* plan);
int dsw_plan_shape(const dsw_plan* plan, int32_t* n_rows, int32_t* n_cols, int64_t* nnz,
                   int32_t* max_row_nnz);
/* Bytes of the sparse operand one pass reads: 8*nnz + 4*(n_rows+1)  (BASELINE.md §4 "nnzB"). */
int64_t dsw_plan_operand_bytes(const dsw_plan* plan);

/* ---------------------------------------------------------------------------------------------
 * Chebyshev graph convolution.  Replaces conv_cheb (modules/layers.py:113-180) + the bias add of
 * ConvCheb.forward (layers.py:365-376):
 *     y[b,v,:] = bias + sum_{k<K} (T_k(L) x_b)[v,:] . W[:,k,:],   T_0=I, T_1=L, T_k=2 L T_{k-1}-T_{k-2}
 * W is the reference's parameter layout [Fin][K][Fout] contiguous; bias [Fout] or NULL.
 * act: 0 = none, 1 = ReLU applied after the bias (ConvBlock.forward, my_models_graph.py:104-118): in the epilogue of the
 * channel mix (TERMS order) or of the last hop (CLENSHAW order).
 * Workspace holds the intermediate planes (K-1 Chebyshev terms of x, or K planes x.W_k).
 * ------------------------------------------------------------------------------------------- */
size_t dsw_cheb_fwd_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K);
int dsw_cheb_fwd(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, const float* W,
                 const float* bias, float* y, int32_t B, int32_t Fin, int32_t Fout, int32_t K,
                 int32_t act, void* workspace, size_t workspace_bytes, void* stream);

/* Only the recurrence: terms[k][b][v][f] for k = 1..K-1 written to `terms` ([K-1][B][V][F]
 * contiguous); term 0 is x itself.  This is the "ChebConv SpMM" stage of BASELINE.json's metric
 * (layers.py:163-169). */
int dsw_cheb_terms(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, float* terms,
                   int32_t B, int32_t F, int32_t K, void* stream);

/* Which evaluation order the forward / backward will use for these channel counts (1 = TERMS,
 * 2 = CLENSHAW; see dsw_cheb.cu): the hops run on the side with fewer channels. */
int dsw_cheb_fwd_algo(int32_t Fin, int32_t Fout, int32_t K);
int dsw_cheb_bwd_algo(int32_t Fin, int32_t Fout, int32_t K);

/* Backward of the convolution (autograd of layers.py:158-177 and of the bias add :375), both
 * gradients in one call so that they can share the Chebyshev terms of dy:
 *     dx_b        = sum_k T_k(L^T) (dy_b . W[:,k,:]^T)                       (dx may be NULL)
 *     dW[f,k,o]   = sum_{b,v} (T_k(L) x_b)[v,f] dy[b,v,o],  dbias[o] = sum_{b,v} dy[b,v,o]   (dW may be NULL)
 * dy is [B][V][Fout] contiguous; dx is written [B][V][Fin] contiguous; dW [Fin][K][Fout] is
 * overwritten; dbias may be NULL.  `saved_terms` = the [K-1][B][V][Fin] Chebyshev terms the forward
 * left at the start of its workspace (only when dsw_cheb_fwd_algo == 1, dsw_cheb_bwd_algo == 2 and
 * sample chunking is off), or NULL.  The reduction over (b, v) is a fixed-order two-pass sum:
 * deterministic. */
size_t dsw_cheb_bwd_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K,
                                    int32_t have_saved_terms);
int dsw_cheb_bwd(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, const float* dy,
                 const float* W, const float* saved_terms, float* dx, float* dW, float* dbias, int32_t B,
                 int32_t Fin, int32_t Fout, int32_t K, void* workspace, size_t workspace_bytes, void* stream);

/* Same with flags.  DSW_BWD_MASK_DX_BY_X: the layer's input x is the output of a ReLU (ConvBlock.forward,
 * my_models_graph.py:104-118, feeding the next ConvBlock of a ResBlock, :205-209), so the gradient that continues upstream is
 * dx * [x > 0] (torch: threshold_backward in a separate pass): the mask is applied by the kernel that writes dx (channel-mix
 * epilogue or last hop).  Needs x (with B > 1: x_sB == V * x_sV, else DSW_ERR_UNSUPPORTED). */
#define DSW_BWD_MASK_DX_BY_X 1
int dsw_cheb_bwd_ex(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV, const float* dy,
                    const float* W, const float* saved_terms, float* dx, float* dW, float* dbias, int32_t B,
                    int32_t Fin, int32_t Fout, int32_t K, int32_t flags, void* workspace, size_t workspace_bytes,
                    void* stream);

/* The two halves on their own (thin wrappers over dsw_cheb_bwd). */
size_t dsw_cheb_bwd_data_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K);
int dsw_cheb_bwd_data(const dsw_plan* lap, const float* dy, const float* W, float* dx, int32_t B,
                      int32_t Fin, int32_t Fout, int32_t K, void* workspace, size_t workspace_bytes,
                      void* stream);
size_t dsw_cheb_bwd_weight_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout, int32_t K);
int dsw_cheb_bwd_weight(const dsw_plan* lap, const float* x, int64_t x_sB, int64_t x_sV,
                        const float* dy, const float* saved_terms, float* dW, float* dbias, int32_t B,
                        int32_t Fin, int32_t Fout, int32_t K, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ---------------------------------------------------------------------------------------------
 * Per-node linear map  y[b,v,:] = x[b,v,:] . Wl^T + bias  (Wl[Fout][Fin], torch.nn.Linear layout) and
 * its gradients: the skip connection of the reference's ResBlock (my_models_graph.py:196-201, :214),
 * on the same tensor-core kernels as the channel mix.  dx / dW / dbias may each be NULL; x must be
 * contiguous for dW.  Deterministic.
 * ------------------------------------------------------------------------------------------- */
size_t dsw_linear_workspace_bytes(int32_t B, int32_t V, int32_t Fin, int32_t Fout);
int dsw_linear_fwd(const float* x, int64_t x_sB, int64_t x_sV, const float* Wl, const float* bias,
                   float* y, int32_t B, int32_t V, int32_t Fin, int32_t Fout, void* workspace,
                   size_t workspace_bytes, void* stream);
int dsw_linear_bwd(const float* x, int64_t x_sB, int64_t x_sV, const float* dy, const float* Wl,
                   float* dx, float* dW, float* dbias, int32_t B, int32_t V, int32_t Fin, int32_t Fout,
                   void* workspace, size_t workspace_bytes, void* stream);
/* Same with the input gradient accumulated:  dx = dy . Wl + dx_addend  (dx_addend [B][V][Fin] contiguous or NULL).
 * A ResBlock's input receives two gradients — the convolution branch's and the skip connection's
 * (my_models_graph.py:205-215; torch adds them in a separate pass): the skip's backward adds the other in its epilogue. */
int dsw_linear_bwd_acc(const float* x, int64_t x_sB, int64_t x_sV, const float* dy, const float* Wl, const float* dx_addend,
                       float* dx, float* dW, float* dbias, int32_t B, int32_t V, int32_t Fin, int32_t Fout, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Sparse remap (interpolation pooling / unpooling).  Replaces RemapBlock.forward
 * (modules/layers.py:956-964) and its autograd:
 *     fwd: y[b,r,:]  = sum_c M[r,c] x[b,c,:]        x [B][n_cols][F] (strided), y [B][n_rows][F]
 *     bwd: dx[b,c,:] = sum_r M[r,c] dy[b,r,:]       dy [B][n_rows][F], dx [B][n_cols][F]
 * ------------------------------------------------------------------------------------------- */
int dsw_spmm_fwd(const dsw_plan* mat, const float* x, int64_t x_sB, int64_t x_sV, float* y,
                 int32_t B, int32_t F, void* stream);
int dsw_spmm_bwd(const dsw_plan* mat, const float* dy, int64_t dy_sB, int64_t dy_sV, float* dx,
                 int32_t B, int32_t F, void* stream);
/* The same two products with an optional addend (indexed like the output rows, element strides a_sB / a_sV, or NULL) and
 * a strided output (row stride >= F):  y = M x + addend,  dx = M^T dy + addend.  They carry the U-Net's skip connection
 * (my_models_graph.py:533,538 `torch.cat((x, x_enc), dim=2)`) without a concatenation pass: the unpool writes its rows
 * straight into one half of the [B][V][2C] tensor the decoder reads (the encoder's last kernel wrote the other half), and
 * the pool's backward adds the skip half of the decoder's input gradient to its own result. */
int dsw_spmm_fwd_ex(const dsw_plan* mat, const float* x, int64_t x_sB, int64_t x_sV, const float* addend, int64_t a_sB,
                    int64_t a_sV, float* y, int64_t y_sB, int64_t y_sV, int32_t B, int32_t F, void* stream);
int dsw_spmm_bwd_ex(const dsw_plan* mat, const float* dy, int64_t dy_sB, int64_t dy_sV, const float* addend, int64_t a_sB,
                    int64_t a_sV, float* dx, int64_t dx_sB, int64_t dx_sV, int32_t B, int32_t F, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Max-value pooling with index output.  Replaces GeneralMaxValPool.forward
 * (modules/layers.py:1043-1083): for coarse row r and column c = f*B + b pick the stored entry j
 * of row r maximising M[r,j]*x[b,j,f] (first maximum wins; NaN counts as maximum, as
 * torch.argmax), output the unweighted x[b,j,f] and the reference's index tensor
 * nnz_ind = int64[2][F*B*Vc]: idx_row[c*Vc + r] = j, idx_col[c*Vc + r] = c.   Bit-exact.
 * bwd = the scatter-add that autograd derives from the final torch.gather (layers.py:1073).
 * ------------------------------------------------------------------------------------------- */
int dsw_maxval_pool_fwd(const dsw_plan* mat, const float* x, int64_t x_sB, int64_t x_sV, float* y,
                        int64_t* idx_row, int64_t* idx_col, int32_t B, int32_t F, void* stream);
int dsw_maxval_pool_bwd(const float* dy, const int64_t* idx_row, float* dx, int32_t B, int32_t V,
                        int32_t Vc, int32_t F, void* stream);

/* Replaces GeneralMaxValUnpool.forward (layers.py:1089-1103): out = zeros[B][V][F];
 * out[b, idx_row[i], f] = x[b, r, f] with i = (f*B+b)*Vc + r and idx_col[i] decoding (f,b).
 * bwd gathers dout at the same positions. */
int dsw_scatter_unpool_fwd(const float* x, const int64_t* idx_row, const int64_t* idx_col,
                           float* out, int32_t B, int32_t V, int32_t Vc, int32_t F, void* stream);
int dsw_scatter_unpool_bwd(const float* dout, const int64_t* idx_row, const int64_t* idx_col,
                           float* dx, int32_t B, int32_t V, int32_t Vc, int32_t F, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Nested-order (HEALPix) pools, window = stride = `kernel` consecutive nodes.  Replace
 * HealpixMaxPool/MaxUnpool (layers.py:784-863) and HealpixAvgPool/AvgUnpool (layers.py:866-941).
 * Max-pool indices are int64 [B][F][V/kernel] holding the fine node position (what
 * F.max_pool1d(return_indices=True) returns on the [B,F,V] view).  Bit-exact.
 * ------------------------------------------------------------------------------------------- */
int dsw_nested_maxpool_fwd(const float* x, int64_t x_sB, int64_t x_sV, float* y, int64_t* idx,
                           int32_t B, int32_t V, int32_t F, int32_t kernel, void* stream);
/* dx[b, idx[b,f,r], f] = dy[b,r,f], zero elsewhere (also the forward of max-unpool). */
int dsw_nested_scatter(const float* src, const int64_t* idx, float* dst, int32_t B, int32_t V,
                       int32_t F, int32_t kernel, void* stream);
/* dst[b,r,f] = src[b, idx[b,f,r], f] (backward of max-unpool). */
int dsw_nested_gather(const float* src, const int64_t* idx, float* dst, int32_t B, int32_t V,
                      int32_t F, int32_t kernel, void* stream);
int dsw_nested_avgpool_fwd(const float* x, int64_t x_sB, int64_t x_sV, float* y, int32_t B,
                           int32_t V, int32_t F, int32_t kernel, void* stream);
/* y[b,v,f] = scale * x[b, v/kernel, f]: scale=1 is avg-unpool fwd (nearest repeat);
 * scale=1/kernel is avg-pool bwd. */
int dsw_nested_repeat(const float* x, int64_t x_sB, int64_t x_sV, float* y, float scale, int32_t B,
                      int32_t V, int32_t F, int32_t kernel, void* stream);
/* y[b,r,f] = sum_{i<kernel} x[b, r*kernel+i, f]  (avg-unpool bwd). */
int dsw_nested_sum(const float* x, int64_t x_sB, int64_t x_sV, float* y, int32_t B, int32_t V,
                   int32_t F, int32_t kernel, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ResBlock tail: y = w * conv_out + skip  (ReZero scale + residual add), n = B*V*F elements, w a
 * device scalar.  Replaces `x_out *= self.rezero_weight; x_out += self.res_connection(x)`
 * (reference modules/my_models_graph.py:211-215) and the passes autograd derives from them.
 * bwd: d_conv_out = w * g (nullable), d_w = sum(g * conv_out) (nullable; needs conv_out and the
 * workspace; fixed-order reduction); the gradient of `skip` is g itself.
 * ------------------------------------------------------------------------------------------- */
/* The whole ResBlock tail in one launch when the skip is a Linear:  y = x . Wl^T + bias + scale[0] * conv_out
 * (`x_out *= self.rezero_weight; x_out += self.res_connection(x)`, my_models_graph.py:196-201, 211-215);
 * conv_out and y are contiguous [B, V, Fout], scale a device scalar.  Workspace as dsw_linear_fwd.  The
 * gradients are dsw_linear_bwd (dy = g) and dsw_rezero_bwd. */
int dsw_linear_rezero_fwd(const float* x, int64_t x_sB, int64_t x_sV, const float* Wl, const float* bias, const float* conv_out,
                          const float* scale, float* y, int32_t B, int32_t V, int32_t Fin, int32_t Fout, void* workspace,
                          size_t workspace_bytes, void* stream);
/* Same, y written with row stride y_ld >= Fout (row n = b*V + v at y + n*y_ld): the encoder's output lands in its half of
 * the skip-concatenation buffer (see dsw_spmm_fwd_ex). */
int dsw_linear_rezero_fwd_ld(const float* x, int64_t x_sB, int64_t x_sV, const float* Wl, const float* bias, const float* conv_out,
                             const float* scale, float* y, int64_t y_ld, int32_t B, int32_t V, int32_t Fin, int32_t Fout,
                             void* workspace, size_t workspace_bytes, void* stream);
int dsw_rezero_fwd(const float* conv_out, const float* skip, const float* w, float* y, int64_t n, void* stream);
size_t dsw_rezero_bwd_workspace_bytes(void);
int dsw_rezero_bwd(const float* g, const float* conv_out, const float* w, float* d_conv_out, float* d_w, void* workspace,
                   size_t workspace_bytes, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Area-weighted MSE loss of the training loop (reference modules/loss.py:118-148 WeightedMSELoss.forward):
 * weighted = weights[v] * (pred - label)^2 on contiguous [B, V, F]; weights [V] or null (= ones).
 * reduction 0 = "mean" (sum / sum(weights) / B / F), 1 = "sum" (the plain sum: the reference's `* len(weights)` acts on
 * the [1, V, 1] view, loss.py:141-144); "none" has its own pair.
 * dsw_wmse_fwd writes the scalar loss and keeps d loss / d sum in the workspace for dsw_wmse_bwd
 * (grad_pred = grad_out[0] * 2 * scale * weights[v] * (pred - label)); fixed-order reductions.
 * ------------------------------------------------------------------------------------------- */
size_t dsw_wmse_workspace_bytes(void);
int dsw_wmse_fwd(const float* pred, const float* label, const float* weights, float* loss, void* workspace, size_t workspace_bytes,
                 int32_t B, int32_t V, int32_t F, int32_t reduction, void* stream);
int dsw_wmse_bwd(const float* pred, const float* label, const float* weights, const void* workspace, const float* grad_out,
                 float* grad_pred, int32_t B, int32_t V, int32_t F, void* stream);
int dsw_wmse_none_fwd(const float* pred, const float* label, const float* weights, float* out, int32_t B, int32_t V, int32_t F,
                      void* stream);
int dsw_wmse_none_bwd(const float* pred, const float* label, const float* weights, const float* grad_out, float* grad_pred, int32_t B,
                      int32_t V, int32_t F, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Introspection used by bench.py / tests.
 * ------------------------------------------------------------------------------------------- */
/* Number of kernels launched by this library since process start (all threads). */
int64_t dsw_launch_count(void);
/* 0 = fp32 CUDA-core channel mix, 1 = tcgen05 split-bf16 (3-term) channel mix when shapes allow. */
int dsw_set_mix_mode(int mode);
int dsw_get_mix_mode(void);
/* Tuning / A-B switches (process-wide; defaults are what production uses).  Values are >= 0. */
enum {
  DSW_OPT_HOP_KERNEL = 0,    /* 0 = auto (bulk-copy-staged tile kernel), 1 = row-block kernel through L1, 2 = plain CSR, 3 = panel-staged L1 tile kernel */
  DSW_OPT_L2_CHUNK_BYTES = 1, /* working-set budget (bytes) of L2-resident sample chunks; 0 / 1 = chunking off (default) */
  DSW_OPT_DEBUG = 2,          /* timing experiments only (results become wrong): 1 = hops skip staging, 2 = hops skip the FMA loop;
                                 dense kernels, bit mask: 16 = no output stores, 32 = no A transfers, 64 = no B transfers,
                                 128 = no bf16 conversion, 256 = no MMAs; 1024 = A/B (results stay right): no L2 prefetch of the mix epilogue's addend;
                                 chain kernel, bit mask: 4 = phase counters, 8 = no stores, 16 = no Z loads, 32 = claim one item further
                                 ahead, 64 = relaxed `done` arrival; 4096 = A/B (results stay right): no two-phase staging */
  DSW_OPT_NO_TMA = 3,         /* 1 = stage tiles with cp.async / register loads instead of tensor-map TMA */
  DSW_OPT_FWD_ALGO = 4,       /* 0 = auto by channel counts, 1 = TERMS (hops on Fin, then mix), 2 = CLENSHAW (mix, then hops on Fout) */
  DSW_OPT_BWD_ALGO = 5,       /* 0 = auto, 1 = TERMS (hops on dy, Fout channels), 2 = CLENSHAW (hops on Fin channels) */
  DSW_OPT_MIX_BN = 6,         /* widest column tile of the tcgen05 channel mix (0 = 256; multiple of 16) */
  DSW_OPT_CONV_MODE = 7,      /* fp32 -> bf16 hi/lo split of the dense kernels: 0 = packed F2FP (default), 1 = one F2F per value, 2 = integer rounding */
  DSW_OPT_HOP_IPC = 8,        /* work items (sample, 64-channel slab) per hop CTA; 0 = heuristic */
  DSW_OPT_HOP_SMALL_F = 9,    /* planes with at most this many channels take the plain CSR hop (0 = default 8) */
  DSW_OPT_HOP_ROWS = 10,      /* extra rows of CTAs per tile in the dynamically scheduled hop (0 = default 2) */
  DSW_OPT_HOP_LPR = 11,       /* lanes per row-block of the tile hop kernel: 0 / 4 = four (default), 8 = eight */
  DSW_OPT_WGRAD_PAIR = 12,    /* 1 = pair CTAs (tcgen05 cta_group::2, 256-row MMAs) in the weight-gradient kernel; measured slower, default off */
  DSW_OPT_HOP_TEAMS = 13,     /* cap on the teams per hop CTA (0 = as many as shared memory holds, up to 5) */
  DSW_OPT_HOP_PREFETCH = 14,  /* the hop kernel L2-prefetches the next item's Z / G rows when it issues its tile transfer (default); 2 = off */
  DSW_OPT_NO_PDL = 15,        /* 1 = launch the hop kernel without programmatic stream serialisation */
  DSW_OPT_PLAN_PERMUTE = 16,  /* locality permutation of the row-block layout at plan creation: 0 = automatic, 1 = never, 2 = always */
  DSW_OPT_NO_CHAIN = 17,      /* 0 = fused persistent chain kernel where it wins (>= one tile per SM), 1 = never (hop-by-hop launches), 2 = wherever it is supported */
  DSW_OPT_CHAIN_L2_BYTES = 18, /* L2 budget (bytes) for the three live planes of a sample group of the chain kernel; 0 = default 64 MiB */
  DSW_OPT_CHAIN_MIN_PASS = 19, /* least number of work items of one hop pass of the chain kernel; 0 = default 2 x teams in flight */
  DSW_OPT_CHAIN_MIN_HOPS = 20, /* chains shorter than this many hops are launched hop by hop; 0 / 1 = every chain is fused */
  DSW_OPT_COUNT = 21
};
/* ---------------------------------------------------------------------------------------------
 * Autoregressive input stacking — the step of the training loop right before model(X) (the
 * reference leaves it to xforecasting.AutoregressiveTraining, scripts_training/train_predict_state.py:392-436, ar_settings
 * of modules/utils_config.py:82-86):  X[b][t][v][:] = [dyn_t[b][v][0..Fd) | bc_t[b][v][0..Fb) | static[v][0..Fs)]  for the T
 * input time slots.  Every slot is a POINTER (an observed state or an earlier prediction), so the shifted history is never
 * materialised: one kernel replaces the shift (cat), the expand of the static fields and the three-way cat.
 * dsw_ar_stack_bwd writes dX's dynamic channels of every slot with ddyn[t] != NULL to that dense [B][V][Fd] buffer.
 * ------------------------------------------------------------------------------------------- */
#define DSW_AR_MAX_SLOTS 8
typedef struct dsw_ar_slots {
  const float* dyn[DSW_AR_MAX_SLOTS];  /* slot t: [B][V][Fd], element strides dyn_sB / dyn_sV, unit feature stride */
  int64_t dyn_sB[DSW_AR_MAX_SLOTS], dyn_sV[DSW_AR_MAX_SLOTS];
  const float* bc[DSW_AR_MAX_SLOTS];   /* slot t: [B][V][Fb] (ignored when Fb == 0) */
  int64_t bc_sB[DSW_AR_MAX_SLOTS], bc_sV[DSW_AR_MAX_SLOTS];
  const float* stat;                   /* [V][Fs] contiguous (ignored when Fs == 0) */
  float* ddyn[DSW_AR_MAX_SLOTS];       /* backward only: gradient buffer of slot t or NULL */
} dsw_ar_slots;
int dsw_ar_stack_fwd(const dsw_ar_slots* slots, float* X, int32_t B, int32_t T, int32_t V, int32_t Fd, int32_t Fb, int32_t Fs, void* stream);
int dsw_ar_stack_bwd(const dsw_ar_slots* slots, const float* dX, int32_t B, int32_t T, int32_t V, int32_t Fd, int32_t Fb, int32_t Fs,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * Device-side construction of the operators that feed the path.
 *
 * dsw_graph_knn_laplacian: symmetrised Gaussian k-NN graph of V unit vectors xyz[V][3] (fp64, device) and its normalised
 * Laplacian L = I - D^-1/2 W D^-1/2 — what the reference obtains from pygsp (modules/models.py:43-46:
 * SphereHealpix(..., k, lap_type="normalized").L): weights exp(-d^2 / (2 sigma^2)), sigma = mean neighbour distance,
 * W = max(W, W^T).  rescale = 1 additionally applies prepare_torch_laplacian (modules/layers.py:82-106): 2 L / lmax - I
 * with lmax = lmax_in if > 0, else a converged, deterministic power iteration x 1.01 (the reference's ARPACK estimate
 * starts from a random vector).  Output: coalesced COO (row-major, ascending columns) into caller-owned device arrays of
 * capacity `cap` >= dsw_graph_nnz_capacity(V, k); *nnz_out / *lmax_out are HOST words (the call synchronises the stream:
 * one-time model construction).  Neighbour ties are broken by the lower node index.
 *
 * dsw_graph_nested_pool: the exact pool ([1/kernel] x kernel per coarse row) / unpool ([1] per fine row) pair of nested
 * orderings (tutorials/interpolation_pooling.ipynb cell 16; the reference computes them with CDO, layers.py:531-581);
 * n_fine entries each.
 * ------------------------------------------------------------------------------------------- */
size_t dsw_graph_workspace_bytes(int32_t V, int32_t k);
int64_t dsw_graph_nnz_capacity(int32_t V, int32_t k);
int dsw_graph_knn_laplacian(const double* xyz, int32_t V, int32_t k, int32_t rescale, double lmax_in, int64_t cap, int64_t* coo_row,
                            int64_t* coo_col, float* coo_val, int64_t* nnz_out, double* lmax_out, void* workspace, size_t workspace_bytes,
                            void* stream);
int dsw_graph_nested_pool(int32_t n_fine, int32_t kernel, int64_t* pool_row, int64_t* pool_col, float* pool_val, int64_t* unpool_row,
                          int64_t* unpool_col, float* unpool_val, void* stream);

/* Tuning only: with DSW_OPT_DEBUG = 4 the hop kernel sums per-phase SM cycles over its teams
 * (issue staging, wait for tile + Z/G, entry loop, stores, item count). */
int dsw_debug_counters(uint64_t* out8, int reset);
/* Tuning only: with DSW_OPT_DEBUG bit 512 the weight-gradient kernel sums role cycles over its CTAs (converters
 * waiting for the TMA, converting, stage count, producer waiting for a free stage, MMA issuer waiting). */
int dsw_debug_dense_counters(uint64_t* out8, int reset);
/* Tuning only: with DSW_OPT_DEBUG bit 2048 the TMA-fed channel mix sums role cycles over its CTAs: [0] producer waiting for
 * a free stage, [1] converters waiting for the TMA, [2] converting, [3] MMA issuer waiting for a free accumulator, [4] ...
 * for operands, [5] epilogue waiting for the accumulator, [6] its TMEM -> shared-memory phase, [7] its store phase,
 * [8] tiles, [9] kernel cycles of CTA 0. */
int dsw_debug_mix_counters(uint64_t* out16, int reset);
/* Tuning only: with DSW_OPT_DEBUG = 4 the fused chain kernel sums per-phase SM cycles: [0..4] over its compute
 * teams (wait for the next item, Z/G loads + wait for the transfers, entry loop, stores, item count), [8..15] over
 * its control warps (metadata, wait for the loop end, issue, wait for the stores, fence + flag, blocking dependency
 * wait, items issued late, item count). */
int dsw_debug_chain_counters(uint64_t* out16, int reset);
int dsw_set_option(int key, int64_t value);
int64_t dsw_get_option(int key);

#ifdef __cplusplus
}
#endif
#endif /* DSW_H_ */
