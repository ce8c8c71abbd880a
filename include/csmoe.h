/* csmoe.h — C ABI of libcsmoe.so: the B200 (sm_100a) kernels behind CompeteSMoE's sparse-MoE layer.
 *
 * Every entry point takes raw device pointers, sizes and a cudaStream_t (passed as void*), returns 0 on success or a
 * negative csmoe_status, never allocates device memory (except csmoe_ep_alloc, which exists to make IPC-exportable
 * exchange buffers), never synchronises the host with the device and is therefore CUDA-graph capturable.  The caller
 * (PyTorch on the host side) owns all other buffers.
 *
 * The reference (Fsoft-AIC/CompeteSMoE) has no native interface; each entry point below names the Python code it
 * replaces (paths relative to the reference root).  See INTEGRATION.md for the binding a maintainer would add.
 */
#ifndef CSMOE_H_
#define CSMOE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSMOE_ABI_VERSION 2

typedef enum csmoe_status {
  CSMOE_OK = 0,
  CSMOE_ERR_ARG = -1,      /* bad argument (null pointer, unsupported size / alignment / dtype) */
  CSMOE_ERR_CUDA = -2,     /* a CUDA runtime call failed; csmoe_last_error() has the text */
  CSMOE_ERR_DRIVER = -3,   /* libcuda entry point (cuTensorMapEncodeTiled) unavailable or failed */
  CSMOE_ERR_UNSUPPORTED = -4
} csmoe_status;

typedef enum csmoe_dtype { CSMOE_F32 = 0, CSMOE_BF16 = 1 } csmoe_dtype;

typedef enum csmoe_act {
  CSMOE_ACT_NONE = 0,
  CSMOE_ACT_RELU = 1,
  CSMOE_ACT_GELU = 2,      /* erf form, torch.nn.GELU() */
  CSMOE_ACT_GELU_TANH = 3, /* transformers ACT2FN["gelu_pytorch_tanh"] (SigLIP MLP) */
  CSMOE_ACT_SILU = 4,
  CSMOE_ACT_SILU_GLU = 5   /* Phi3MLP: cols [0,F) = gate, [F,2F) = up; h = up * silu(gate) */
} csmoe_act;

/* Row tile of every grouped operand: expert segments in the permuted ("expert-major") row space start at multiples of
 * this, so that one GEMM tile never straddles two experts. */
#define CSMOE_ROW_TILE 128

int csmoe_abi_version(void);
/* Text of the last failure on the calling thread ("" if none). */
const char* csmoe_last_error(void);
/* 1 when the current device is compute capability 10.x, 0 otherwise, <0 on error. */
int csmoe_device_supported(void);

/* ------------------------------------------------------------------------------------------------ routing metadata
 * Replaces cvmm_prepare_sel2 (moe_pretrain_model/layers/cvmm.py:580-592: flatten -> sort -> index maps) and the
 * E x torch.where of MoeLayer.compute_moe (moe_model/model/moe/moe.py:189-191).
 *
 * In : sel[T*K] int32 expert id of slot j = t*K + k.
 * Out: counts[E], offsets[E+1] (exclusive scan of counts = the reference's sorted segment boundaries),
 *      pad_offsets[E+1] (segment starts rounded up to row_tile = 128 or 256; 256 enables the CTA-pair GEMM),
 *      sorted_sel[T*K], sort_index[T*K] (stable argsort of sel: the reference's ssel / out_index; in_index = /K),
 *      slot_to_row[T*K] (row of slot j in the padded expert-major space), row_to_slot[row_cap] (-1 for padding rows),
 *      tile_expert[row_cap/CSMOE_ROW_TILE] (expert owning each row tile, -1 past the end).
 * row_cap must be >= csmoe_route_row_cap(T*K, E, row_tile).  Any output pointer except counts/offsets/pad_offsets may be NULL.
 * workspace: csmoe_route_workspace_bytes(T*K, E) bytes. */
int64_t csmoe_route_row_cap(int64_t n_slots, int32_t num_experts, int32_t row_tile);
int64_t csmoe_route_workspace_bytes(int64_t n_slots, int32_t num_experts);
int csmoe_route_build(const int32_t* sel, int64_t n_slots, int32_t num_experts, int32_t row_tile, int64_t row_cap, int32_t* counts,
                      int32_t* offsets, int32_t* pad_offsets, int32_t* sorted_sel, int64_t* sort_index,
                      int32_t* slot_to_row, int32_t* row_to_slot, int32_t* tile_expert, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------ router
 * Replaces router_policy + topk_expert (moe_model/model/moe/competesmoe.py:301-320, moe.py:113-132;
 * moe_pretrain_model/layers/moe/competesmoe.py:465-490): logits = x @ Wg^T (fp32 accumulate, rounded to `x_dtype`),
 * p = softmax_fp32(logits), (w, idx) = topk(p, K) (descending, ties -> lowest index), w /= round_to(renorm_dtype)(sum w):
 * the reference's `.to(x.dtype)` on the denominator refers to the LAYER INPUT, which is the activation dtype for a bf16
 * model and fp32 for fp32 inputs run under autocast (then x_dtype = bf16, renorm_dtype = fp32).
 * x[T,D] and wg[E,D] share x_dtype (bf16 or fp32).  Outputs: logits[T,E] (x_dtype), probs[T,E] fp32, topk_w[T,K] fp32,
 * topk_idx[T,K] int32.  E <= 256 (two experts per lane up to 64, four / eight per lane above), K <= 8. */
int csmoe_router_fwd(const void* x, const void* wg, int32_t x_dtype, int64_t T, int32_t D, int32_t E, int32_t K,
                     int32_t renorm_dtype, void* logits, float* probs, float* topk_w, int32_t* topk_idx, void* stream);

/* The softmax / top-k / renormalisation half of csmoe_router_fwd for logits [T, E] that were produced elsewhere (the
 * gate GEMM on the tensor cores when E is large: csmoe_grouped_gemm with one "expert" = the gate matrix). */
int csmoe_router_from_logits(const void* logits, int32_t dtype, int64_t T, int32_t E, int32_t K, int32_t renorm_dtype,
                             float* probs, float* topk_w, int32_t* topk_idx, void* stream);
/* Top-k over given fp32 scores[T,E] (competition step: affinity scores; moe_model/.../competesmoe.py:249-254).
 * mode 0: w = topk values; mode 1: w = sigmoid(topk values) (norm_sigmoid).  Then w /= round_to(dtype)(sum w). */
int csmoe_topk_renorm(const float* scores, int64_t T, int32_t E, int32_t K, int32_t mode, int32_t round_dtype,
                      float* topk_w, int32_t* topk_idx, void* stream);

/* Router-step auxiliary losses in one pass (moe_model/model/moe/moe.py:71-110,214-226):
 *   losses[0] = balance = E^2 * mean_{b,e}( mean_n probs[b,n,e] * mean_n [top-1(b,n) == e] )
 *   losses[1] = z-loss  = mean_t logsumexp(logits_t)^2
 * Also returns psum[B,E] = sum_n probs, cnt[B,E] = top-1 counts and lse[T] (may be NULL) for the backward pass.
 * Two-stage, fixed-order reduction (deterministic).  workspace: csmoe_router_aux_workspace_bytes(B, N, E). */
int64_t csmoe_router_aux_workspace_bytes(int64_t B, int64_t N, int32_t E);
int csmoe_router_aux_fwd(const void* logits, int32_t dtype, const float* probs, const int32_t* topk_idx, int64_t B,
                         int64_t N, int32_t E, int32_t K, float* psum, float* cnt, float* lse, float* losses,
                         void* workspace, void* stream);

/* Router backward, fused: gradient w.r.t. the gate logits from (optional, NULL = absent) the routing-weight gradient
 * dtw[T,K], an incoming dprobs[T,E] (e.g. router-distillation MSE), an incoming dlogits[T,E], and the balance / z
 * losses (g_losses[2] = d loss / d balance, d loss / d z on the device; cnt and lse from csmoe_router_aux_fwd);
 * then dx[T,D] = dl . Wg (may be NULL) and dWg[E,D] = dl^T . x (may be NULL; deterministic two-stage reduction).
 * The routing-weight term follows autograd through `w / sum(w).to(dtype)`: numerator gradient dtw / r in fp32, denominator
 * gradient -sum_k dtw_k (w_k / r) rounded to renorm_dtype (r = the rounded denominator of the forward pass).
 * dl[T,E] (fp32 storage, values rounded to x_dtype) is an output.  workspace: csmoe_router_bwd_workspace_bytes. */
int64_t csmoe_router_bwd_workspace_bytes(int64_t T, int32_t D, int32_t E);
int csmoe_router_bwd(const void* x, const void* wg, int32_t x_dtype, const float* probs, const float* topk_w,
                     const int32_t* topk_idx, const float* dtw, const float* dprobs, const float* dlogits,
                     const float* lse, const float* cnt, const float* g_losses, int64_t B, int64_t N, int32_t D, int32_t E,
                     int32_t K, int32_t renorm_dtype, float* dl, void* dx, void* dwg, int32_t wg_dtype, void* workspace,
                     void* stream);

/* ------------------------------------------------------------------------------------------------ permutation
 * Gather rows of src[T, D] into the padded expert-major space: dst[row] = scale(row) * src[row_to_slot[row] / K],
 * zero for padding rows.  scale = slot_w[slot] when slot_w != NULL (combine backward), else 1.
 * Replaces x[batch_idx, token_idx] (moe.py:199) and the gathered A loads of cvmm_kernel (cvmm.py:114-118). */
int csmoe_gather_rows(const void* src, int32_t dtype, int64_t T, int32_t D, int32_t K, const int32_t* row_to_slot,
                      int64_t row_cap, const float* slot_w, void* dst, void* stream);

/* Combine: out[t] = sum_k w[t,k] * y[slot_to_row[t*K+k]]  in ascending-expert order (deterministic), replacing the
 * in-place `results[b,t] += w * out` loop (moe.py:204) and the bmm reduce of CVMM.forward (cvmm.py:481-483).
 * round_each != 0 reproduces the reference's rounding of the running sum to `dtype` after every expert (moe.py:204).
 * sel[T*K] gives the expert of each slot (ordering key).  bias (may be NULL): o_bias[D] added at the end. */
int csmoe_combine_fwd(const void* y, int32_t dtype, int64_t T, int32_t D, int32_t K, const int32_t* slot_to_row,
                      const int32_t* sel, const float* w, int32_t round_each, void* out, void* stream);
/* dw[t,k] = <dout[t], y[row(t,k)]> (fp32). */
int csmoe_combine_bwd_w(const void* y, const void* dout, int32_t dtype, int64_t T, int32_t D, int32_t K,
                        const int32_t* slot_to_row, float* dw, void* stream);
/* Un-permute with reduction: dx[t] (+)= sum_k g[slot_to_row[t*K+k]]; accumulate != 0 adds to the existing dx. */
int csmoe_scatter_reduce(const void* g, int32_t dtype, int64_t T, int32_t D, int32_t K, const int32_t* slot_to_row,
                         int32_t accumulate, void* dx, void* stream);

/* ------------------------------------------------------------------------------------------------ grouped GEMM
 * tcgen05/TMEM grouped GEMM fed by TMA; bf16 operands, fp32 accumulation.
 * Replaces cvmm_kernel / cvmm_backward_kernel3 (moe_pretrain_model/layers/cvmm.py:61-168, 194-345) and the per-expert
 * nn.Linear calls of compute_moe / competition_policy (moe_model/model/moe/moe.py:196-204, competesmoe.py:240-245).
 *
 * mode ROWS   : C[row, :] = A[row, :] . B[expert(row)]   for row tiles of the padded expert-major space
 *               (forward and dgrad).  A is [m, k] row-major.  b_layout 0: B[e] is [n, k] row-major (nn.Linear weight,
 *               forward); b_layout 1: B[e] is [k, n] row-major (sigma-MoE keys/values forward, nn.Linear dgrad).
 * mode REDUCE : C[e] = A[rows of e, :]^T . B[rows of e, :]   (wgrad).  A is [rows, m], B is [rows, n] row-major,
 *               C[e] is [m, n].
 * dense != 0  : every expert processes the same dense_rows rows (competition step, all experts on all tokens):
 *               ROWS: A row tile = tile % (dense_rows/128) (+ e*a_expert_rows), C row = e*dense_rows + ...;
 *               REDUCE: A rows of e start at e*a_expert_rows, B rows at e*b_expert_rows (0 = shared operand).
 * Epilogue    : + bias[e][n] (optional), activation (optional, ROWS only); when preact != NULL the pre-activation value
 *               is also stored (same layout/dtype as C).  act = SILU_GLU (ROWS, b_layout 0): B[e] is [2F, k] (gate rows
 *               then up rows), n = 2F, preact is [m, 2F], C = up * silu(gate) is [m, F]; each output tile computes 128
 *               gate and the matching 128 up columns so the product never leaves the epilogue.
 *               act_bwd != NONE (ROWS): C = acc * act'(aux) with aux the saved pre-activation (fused activation
 *               backward of the dgrad GEMM); act_bwd = SILU_GLU: acc is dh [m, n = F], aux is [m, 2F], C is [m, 2F].
 * All leading dimensions are in elements and must make rows 16-byte aligned; n % 8 == 0. */
typedef enum csmoe_gemm_mode { CSMOE_GEMM_ROWS = 0, CSMOE_GEMM_REDUCE = 1 } csmoe_gemm_mode;

typedef struct csmoe_gemm_args {
  int32_t mode;
  int32_t b_layout;
  int32_t num_experts;
  int32_t dense;
  int64_t m, n, k;
  int64_t dense_rows;
  const void* a;
  int64_t lda;
  int64_t a_expert_rows;
  const void* b;
  int64_t ldb;
  int64_t b_expert_stride; /* ROWS: elements between consecutive experts' weights. REDUCE+dense: rows between experts */
  void* c;
  int64_t ldc;
  int64_t c_expert_stride; /* REDUCE: elements between consecutive experts' outputs */
  int32_t c_dtype;
  int32_t act;
  const void* bias;        /* [num_experts, n] or NULL */
  int32_t bias_dtype;
  int32_t accumulate;      /* REDUCE: C[e] += result (fp32 C only) */
  void* preact;
  int64_t ldpre;
  const int32_t* tile_expert; /* ROWS, !dense: [m / 128] */
  const int32_t* pad_offsets; /* REDUCE, !dense: [num_experts + 1] */
  int32_t max_ctas;        /* 0 = one per SM */
  int32_t act_bwd;         /* ROWS: C = acc * act'(aux) (dgrad through the activation; SILU_GLU: C is [m, 2n]) */
  const void* aux;         /* saved pre-activation z for act_bwd: [m, n] (SILU_GLU: [m, 2n]), bf16 */
  int64_t ldaux;
  int32_t row_tile;        /* ROWS, !dense: the row_tile the routing maps were built with (128 or 256); 256 lets the
                              CTA-pair (cta_group::2, 256 x 256 tile) kernel run */
  int32_t sum_experts;     /* ROWS + dense + a_expert_rows > 0: C[dense_rows, n] = sum over experts e of
                              A[e*a_expert_rows + row, :] . B[e] -- one launch whose k loop also runs over the experts
                              (dgrad of the competition step's dense pass w.r.t. the shared input) */
  const uint64_t* c_rows;  /* ROWS, plain epilogue: when non-NULL, output row r is stored at address c_rows[r] (0 = row
                              skipped) instead of c + r*ldc -- the expert-parallel return path: the down projection
                              writes each row straight into the source rank's buffer (peer memory).  c may be NULL. */
  float* rowsum;           /* ROWS, plain epilogue, may be NULL: rowsum[row * ceil(n/64) + g] = sum over output columns
                              [64g, 64g+64) of softplus(C[row, col]) (of the value as stored) -- the competition step's
                              neural-response score reduced in the epilogue of the experts' down projection
                              (moe_model/.../competesmoe.py:243: mean(softplus(out_i), -1)); csmoe_affinity_from_rowsum
                              finishes the mean.  Rows of skipped tiles are not written. */
  int32_t rowsum_round;    /* 1: round every softplus to bf16 before summing (eager bf16 reference arithmetic) */
  int32_t bias_after_round; /* 1: the accumulator is rounded to bf16 BEFORE the bias is added -- the pretrain plugin's
                              `scores = cvmm(...)` (bf16) `+ self.bias[...]` (fp32), moe.py:397-401; 0: bias added to the fp32
                              accumulator, one rounding (nn.Linear with bias, the multimodal experts) */
} csmoe_gemm_args;

int csmoe_grouped_gemm(const csmoe_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------ activations
 * Elementwise forward/backward used where the activation is not fused into a GEMM epilogue.
 * fwd: h = act(z).  GLU: z is [rows, 2F], h is [rows, F].   bwd: dz = dh * act'(z).
 * tile_expert (may be NULL): the routing's [rows / 128] tile map; 128-row tiles with a negative entry hold no routed
 * rows and are skipped (their outputs are left untouched) -- the padded row space is sized for the worst case. */
int csmoe_act_fwd(const void* z, int32_t dtype, int64_t rows, int64_t cols, int64_t ldz, int32_t act, void* h,
                  int64_t ldh, const int32_t* tile_expert, void* stream);
int csmoe_act_bwd(const void* z, const void* dh, int32_t dtype, int64_t rows, int64_t cols, int64_t ldz, int64_t ldh,
                  int32_t act, void* dz, const int32_t* tile_expert, void* stream);
/* Column sums of g[rows of e, n] per expert -> dbias[e, n] (fp32 accumulate, written as `out_dtype`).  workspace (may
 * be NULL; csmoe_bias_grad_workspace_bytes(n, E) bytes): lets the kernel split every expert's rows over several CTAs
 * and add the partial sums in a fixed order (deterministic); without it one CTA column reduces all rows of an expert. */
int64_t csmoe_bias_grad_workspace_bytes(int32_t n, int32_t num_experts);
int csmoe_bias_grad(const void* g, int32_t dtype, int64_t ldg, int32_t n, int32_t num_experts,
                    const int32_t* pad_offsets, int32_t dense, int64_t dense_rows, void* dbias, int32_t out_dtype,
                    void* workspace, void* stream);
/* Activation backward and bias gradient of the first projection in one pass: dz[r, :] = dh[r, :] * act'(z[r, :])
 * (rounded to the activation dtype, stored) and dbias[e, :] = sum over expert e's rows of dz -- the two steps autograd
 * runs after `expert.fc1` / `experts[i][0]` (moe_model/model/moe/moe.py:196-204 backward).  Same row ranges, workspace
 * and determinism as csmoe_bias_grad; act = RELU / GELU / GELU_TANH / SILU. */
int csmoe_act_bwd_bias(const void* z, const void* dh, int32_t dtype, int64_t ldz, int64_t ldh, int32_t n,
                       int32_t num_experts, const int32_t* pad_offsets, int32_t dense, int64_t dense_rows, int32_t act,
                       void* dz, void* dbias, int32_t out_dtype, void* workspace, void* stream);
/* dst = (bf16) src, n elements. */
int csmoe_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);
/* hi + mid + lo == src to 24 bits, each term bf16: operands of the fp32-accurate path, in which one fp32 GEMM is six bf16
 * tensor-core products (hi.hi, hi.mid, mid.hi, mid.mid, hi.lo, lo.hi) accumulated into an fp32 C with
 * csmoe_gemm_args.accumulate -- fp32 callers outside autocast get fp32 results (the reference's fp32 cvmm is IEEE FMA,
 * layers/cvmm.py:395 allow_tf32=False; BASELINE north_star: fp32 rtol 1e-4). */
int csmoe_split_f32_bf16x3(const float* src, void* hi, void* mid, void* lo, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------ competition
 * Neural-response score (moe_model/.../competesmoe.py:240-243; moe_pretrain_model/.../competesmoe.py:399-403):
 * aff[t, e] = mean_d softplus(y[e, t, d]) over dense expert outputs y[E, t_pad, D] (the grouped GEMM's dense layout).
 * round_dtype = CSMOE_BF16 reproduces eager bf16 arithmetic (each softplus and the mean rounded to bf16, as the
 * multimodal reference computes it in x.dtype); CSMOE_F32 keeps fp32 (the pretrain reference under autocast, where
 * softplus and mean run in fp32).  aff is always stored as fp32. */
int csmoe_affinity_fwd(const void* y, int32_t dtype, int32_t E, int64_t T, int64_t t_pad, int32_t D,
                       int32_t round_dtype, float* aff, void* stream);
/* aff[t, e] = (sum_g rowsum[(e * t_pad + t) * groups + g]) / D, rounded to bf16 when round_dtype = CSMOE_BF16: finishes the
 * score from the row sums the grouped GEMM's epilogue produced (csmoe_gemm_args.rowsum). */
int csmoe_affinity_from_rowsum(const float* rowsum, int32_t groups, int32_t E, int64_t T, int64_t t_pad, int32_t D,
                               int32_t round_dtype, float* aff, void* stream);
/* dy[e,t,d] (+)= daff[t,e] * sigmoid(y[e,t,d]) / D. accumulate: add into an existing dy. */
int csmoe_affinity_bwd(const void* y, const float* daff, int32_t dtype, int32_t E, int64_t T, int64_t t_pad, int32_t D,
                       int32_t accumulate, void* dy, void* stream);

/* Diversity loss of the selected experts' outputs (moe_model/.../competesmoe.py:180-218; moe_pretrain_model/.../
 * competesmoe.py:330-372): per token the K x K cosine-similarity matrix of rows y[sel[t,k], t, :] (fp32 math,
 * F.normalize eps 1e-12); loss[0] = sum of off-diagonal entries / (T*K*K), reduced in a fixed order.
 * Saved for backward: inv_norm [T,K], sim [T,K,K]; partial [T] is scratch. */
int csmoe_diversity_fwd(const void* y, int32_t dtype, int64_t T, int64_t t_pad, int32_t D, int32_t K, const int32_t* sel,
                        float* inv_norm, float* sim, float* partial, float* loss, void* stream);
/* Whole gradient of the dense expert outputs y[E, t_pad, D] of a competition step in one pass (replaces the autograd
 * of competition_policy + compute_moe + experts_diversity_loss, moe_model/.../competesmoe.py:219-259,:371-374):
 *   dy[e,t,:] = daff[t,e] * sigmoid(y) / D + [e == sel[t,k]] * (w[t,k] * dout[t,:] + d(diversity)/dy * g_div[0]).
 * daff [T,E], (w [T,K], dout [T,D] of y's dtype), (inv_norm, sim, g_div = device scalar) may each be NULL = that
 * term is absent.  Rows t in [T, t_pad) of dy are zeroed. */
int csmoe_compete_bwd(const void* y, int32_t dtype, int32_t E, int64_t T, int64_t t_pad, int32_t D, int32_t K,
                      const float* daff, const int32_t* sel, const float* w, const void* dout, const float* inv_norm,
                      const float* sim, const float* g_div, void* dy, void* stream);

/* ------------------------------------------------------------------------------------------------ fused sigma-MoE FFN
 * The pretrain plugin's expert path -- cvmm(x, sel, keys) -> (+bias) -> relu -> cvmm(., sel, values), moe_pretrain_model/
 * layers/moe/competesmoe.py:510-522 with layers/moe/moe.py:397-416 -- as one kernel per direction for expert size
 * H = 128 (every sweep of the reference; transformer_lm_mixin.py:33).  Rows live in the padded expert-major row space
 * of csmoe_route_build (row_to_slot / tile_expert / pad_offsets); the token rows are read straight from the
 * token-major tensor with TMA gather4 loads (row of slot j = j / slots_per_row), so the K-fold expanded copy of the
 * tokens is never written and the [rows, H] hidden activations stay on chip between the two MMAs.  All tensors bf16.
 *   csmoe_sigma_ffn_supported(D, H, Dout): 1 if this path handles the shape (H == 128, D % 64 == 0, Dout % 128 == 0).
 *   fwd:  h [row_cap, H] = relu(gather(x) . keys[e] + bias[e]) (saved for backward, zero on padding rows);
 *         y [row_cap, Dout] = h . values[e].   xp (may be NULL) = a pre-gathered [row_cap, D] copy to load instead.
 *   bwd:  dh = gather(dout) . values[e]^T;  dw_part[half][slot] = partial <h, dh> (sum the two halves: d routing weight);
 *         dz [row_cap, H] = slot_w * dh * [h > 0];  hw [row_cap, H] = slot_w * h;  dxr [row_cap, D] = dz . keys[e]^T.
 *         dyp (may be NULL) = pre-gathered UNWEIGHTED dout rows.  Needs D % 128 == 0 as well.
 *   wgrad: c[e] = a[rows of e]^T . gather(g)[rows of e] with a [row_cap, H], g token-major [T, N], N % 128 == 0:
 *         transpose = 0 -> c [E, H, N] (dvalues = hw^T . gather(dout));  1 -> c [E, N, H] (dkeys = (dz^T . gather(x))^T).
 *         fp32 accumulation over the expert's rows inside one CTA: deterministic (cvmm.py:194-345 uses fp32 atomics). */
int csmoe_sigma_ffn_supported(int64_t D, int32_t H, int64_t Dout);
/* tuning aid: per-CTA wait-cycle counters of csmoe_sigma_wgrad are written to buf (8 x uint64 per CTA); NULL switches it off */
int csmoe_sigma_set_stats(void* buf);
int csmoe_sigma_ffn_fwd(const void* x, int64_t T, int32_t D, int32_t Dout, int32_t E, const void* keys, const void* values,
                        const void* bias, int32_t bias_dtype, const int32_t* row_to_slot, const int32_t* tile_expert,
                        int64_t row_cap, int32_t slots_per_row, const void* xp, void* h, void* y, void* stream);
int csmoe_sigma_ffn_bwd(const void* dout, int64_t T, int32_t D, int32_t Dout, int32_t E, const void* keys,
                        const void* values, const int32_t* row_to_slot, const int32_t* tile_expert, int64_t row_cap,
                        int32_t slots_per_row, const float* slot_w, int64_t n_slots, const void* h, const void* dyp,
                        void* dz, void* hw, void* dxr, float* dw_part, void* stream);
int csmoe_sigma_wgrad(const void* a, const void* g, int64_t T, int32_t N, int32_t E, const int32_t* row_to_slot,
                      const int32_t* pad_offsets, int64_t row_cap, int32_t slots_per_row, int32_t transpose, void* c,
                      int32_t c_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------ losses
 * Competition-step losses in one pass over [T, E] (T = B * N tokens, batch-major) + a fixed-order reduction.  Replaces
 * F.softmax(affinity), the router-distillation MSE and its in_topk / hybrid / tribrid gathers, balanceloss(aff_idx,
 * aff_softmax) and entropy_balance(aff_softmax): moe_model/model/moe/competesmoe.py:322-335,350-371, moe.py:90-110;
 * moe_pretrain_model/layers/moe/competesmoe.py:541-593, layers/moe/moe.py:323-332.
 *   in : p [T,E] gate softmax, aff [T,E] affinity scores, aff_idx [T,K] competition top-k, gate_idx [T,K] router top-k
 *        (may be NULL: losses[2] = 0)
 *   out: q [T,E] = softmax(aff);  losses[5]:
 *        [0] mean_{t,e}(p - q)^2                         F.mse_loss(gate_softmax, affinity_softmax)
 *        [1] mean_{t,k}(p - q)^2 at aff_idx              F.mse_loss of the gathered top-k columns (hybrid / in_topk)
 *        [2] the same at gate_idx                         (tribrid)
 *        [3] mean_{b,e}(mean_n q * mean_n onehot(aff_idx[..., 0])) * E^2      balanceloss on the affinity (multimodal)
 *        [4] mean_b sum_e m log m,  m = mean_n softmax(q)                      entropy_balance(aff_softmax) (pretrain)
 *        colq / cnt / colr [B,E]: per-batch column sums of q, top-1 counts and column sums of softmax(q) (for backward).
 * workspace: csmoe_losses_workspace_bytes(B, N, E) bytes.  E <= 256, K <= 8, B <= 65535. */
int64_t csmoe_losses_workspace_bytes(int64_t B, int64_t N, int32_t E);
int csmoe_losses_fwd(const float* p, const float* aff, const int32_t* aff_idx, const int32_t* gate_idx, int64_t B,
                     int64_t N, int32_t E, int32_t K, float* q, float* colq, float* cnt, float* colr, float* losses,
                     void* workspace, void* stream);
/* g[5] (device) = d(total)/d(losses[i]).  dp [T,E]: gradient of the MSE terms w.r.t. the gate softmax (q is detached
 * there, as in the reference); daff [T,E]: gradient of terms [3] and [4] w.r.t. the affinity scores, through q. */
int csmoe_losses_bwd(const float* p, const float* q, const int32_t* aff_idx, const int32_t* gate_idx, const float* cnt,
                     const float* colr, const float* g, int64_t B, int64_t N, int32_t E, int32_t K, float* dp,
                     float* daff, void* stream);
/* Router-step regulariser of the pretrain layer, entropy_balance(gate_logits) (layers/moe/moe.py:323-332), from the
 * probabilities the router kernel already produced: loss[0] = mean_b sum_e m log m with m[b,e] = mean_n probs[b,n,e];
 * colr [B,E] = sum_n probs.  Backward: dprobs[t,e] = g[0] * (log m[b,e] + 1) / (B * N).  Same workspace query. */
int csmoe_entropy_balance_fwd(const float* probs, int64_t B, int64_t N, int32_t E, float* colr, float* loss,
                              void* workspace, void* stream);
int csmoe_entropy_balance_bwd(const float* colr, const float* g, int64_t B, int64_t N, int32_t E, float* dprobs,
                              void* stream);
/* Backward of csmoe_topk_renorm: w = v / sum(v) with v = scores (or sigmoid(scores)) at idx; dscores [T,E] is written
 * (accumulate = 0) or added to (1).  The renormalised weights stay attached to the scores in the reference
 * (competesmoe.py:249-254), so this runs in every competition step. */
int csmoe_topk_renorm_bwd(const float* scores, const float* w, const int32_t* idx, const float* dw, int64_t T, int32_t E,
                          int32_t K, int32_t sigmoid, int32_t accumulate, float* dscores, void* stream);
/* rows[t*K + k] = idx[t,k] * t_pad + t: where the selected experts' rows sit in the dense outputs y[E, t_pad, D]. */
int csmoe_dense_rows(const int32_t* idx, int64_t T, int32_t K, int64_t t_pad, int32_t* rows, void* stream);

/* ------------------------------------------------------------------------------------------------ block tail
 * The steps either side of the layer in the reference's pre-LN transformer block
 * (moe_pretrain_model/layers/transformer/relative_moe_transformer.py:150-159):
 *     src2 = norm2(src);  src3 = pkm(src2, id_layer);  src = src + dropout(src3)
 * csmoe_layernorm_fwd: y = LayerNorm(x) with fp32 statistics, written in y_dtype (fp32 -> bf16: the autocast cast that
 *   follows the LayerNorm is folded in); mean / rstd [T] are saved for backward.
 * csmoe_layernorm_bwd: dx (x's dtype), dgamma / dbeta [D] fp32 (two-stage, fixed order); workspace from the query.
 * csmoe_combine_residual_fwd: csmoe_combine_fwd whose epilogue computes out = residual + dropout(combined rows), the
 *   combined value rounded to the row dtype first (it is the layer's bf16 output in the reference), out in res_dtype.
 * csmoe_residual_dropout_fwd: the same tail on a finished layer output v [T, D].
 * csmoe_dropout_bwd: dv = keep ? g * 1/(1-p) : 0 -- the mask is regenerated from (seed, element index): Philox4x32-10,
 *   counter = element index / 4.  p = 0 switches dropout off (seed ignored). */
int csmoe_layernorm_fwd(const void* x, int32_t x_dtype, int64_t T, int32_t D, const float* gamma, const float* beta, float eps,
                        void* y, int32_t y_dtype, float* mean, float* rstd, void* stream);
int64_t csmoe_layernorm_bwd_workspace_bytes(int64_t T, int32_t D);
int csmoe_layernorm_bwd(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* mean, const float* rstd,
                        const float* gamma, int64_t T, int32_t D, void* dx, float* dgamma, float* dbeta, void* workspace,
                        void* stream);
int csmoe_combine_residual_fwd(const void* y, int32_t dtype, int64_t T, int32_t D, int32_t K, const int32_t* slot_to_row,
                               const int32_t* sel, const float* w, int32_t flags, const void* residual, int32_t res_dtype,
                               float p, uint64_t seed, void* out, void* stream);
int csmoe_residual_dropout_fwd(const void* v, int32_t v_dtype, const void* residual, int32_t res_dtype, int64_t T, int32_t D,
                               float p, uint64_t seed, void* out, void* stream);
int csmoe_dropout_bwd(const void* g, int32_t g_dtype, int64_t n, float p, uint64_t seed, void* dv, int32_t v_dtype,
                      void* stream);

/* ------------------------------------------------------------------------------------------------ expert parallelism
 * One process per GPU; rank r of P owns experts [r*E/P, (r+1)*E/P).  Exchange buffers are allocated by the library
 * (cudaMalloc, so that they can be exported with CUDA IPC) and mapped into every peer once; after that dispatch and
 * return are kernels that store straight into peer HBM over NVLink.  The reference has no counterpart: it is
 * data-parallel only (moe_pretrain_model/framework/task/simple_task.py:403-413, moe_model/train/train.py:1474-1480).
 * "peer arrays" below are HOST arrays of P device pointers, entry r = the same buffer on rank r (own rank: local ptr).
 * These are the only entry points that allocate device memory (csmoe_ep_alloc) -- everything else is caller-owned. */
#define CSMOE_EP_MAX_RANKS 16
int csmoe_ep_ipc_handle_bytes(void);
/* cudaMalloc + zero-fill `bytes`; writes csmoe_ep_ipc_handle_bytes() bytes of IPC handle to handle_out (may be NULL). */
int csmoe_ep_alloc(int64_t bytes, void** ptr, void* handle_out);
int csmoe_ep_open(const void* handle, void** ptr);  /* map a peer's allocation (enables peer access lazily) */
int csmoe_ep_close(void* ptr);
int csmoe_ep_free(void* ptr);
/* Cross-rank barrier on the stream: flags = peer array of int32[P] flag vectors, epoch = this rank's device counter.
 * Release/acquire at system scope: everything the calling rank stored to peer memory before the barrier is visible to
 * the peers' kernels after it.  Every rank must issue the same sequence of barrier-containing calls. */
int csmoe_ep_barrier(const void* const* flags, int32_t* epoch, int32_t rank, int32_t P, void* stream);
/* Publish counts[E] (this rank's rows per expert, from csmoe_route_build) into every peer's counts_all[P][E], barrier,
 * then derive: dest_base[E] = row in the owner's receive space where this rank's rows for expert e start;
 * recv_counts[E/P], recv_pad_offsets[E/P + 1], tile_expert[row_cap/128] = padded expert-major layout (segments
 * aligned to row_tile, rows ordered by source rank, then source order) of the rows this rank receives.
 * row_cap >= csmoe_route_row_cap(P * n_slots_max, E/P, row_tile) is the static capacity of the receive buffers. */
int csmoe_ep_exchange_plan(const int32_t* counts, const void* const* counts_all, const void* const* flags, int32_t* epoch,
                           int32_t rank, int32_t P, int32_t E, int32_t row_tile, int64_t row_cap, int32_t* dest_base,
                           int32_t* recv_counts, int32_t* recv_pad_offsets, int32_t* tile_expert, void* stream);
/* Permute + send: row src[j / K] (scaled by slot_w[j] when given) of slot j goes to row
 * dest_base[sel[j]] + (slot_to_row[j] - pad_offsets[sel[j]]) of peer sel[j] / experts_per_rank's recv buffer [row_cap, D];
 * tags (peer array of int64[row_cap], may be NULL) receives (rank << 32 | j) for the return trip. */
int csmoe_ep_dispatch(const void* src, int32_t dtype, int32_t D, int32_t K, int64_t n_slots, const int32_t* sel,
                      const int32_t* slot_to_row, const int32_t* pad_offsets, const int32_t* dest_base,
                      int32_t experts_per_rank, const float* slot_w, const void* const* recv, const void* const* tags,
                      int32_t rank, int32_t P, void* stream);
/* Receiver side, after the barrier that follows dispatch: c_rows[r] (may be NULL) = address of row `slot` of peer
 * `src`'s return buffer ret[src] ([n_slots, ret_ld] of `dtype`) for valid rows, 0 for padding; and the padding rows of
 * recv [row_cap, D] (may be NULL) are zeroed (the wgrad GEMM contracts over whole padded segments). */
int csmoe_ep_row_ptrs(const int64_t* tags, const int32_t* tile_expert, const int32_t* recv_counts,
                      const int32_t* recv_pad_offsets, int32_t experts_per_rank, int64_t row_cap, const void* const* ret,
                      int64_t ret_ld, int32_t dtype, int32_t P, uint64_t* c_rows, void* recv, int32_t D, void* stream);
/* Stand-alone return transfer (the grouped GEMM's c_rows epilogue does the same inside the down projection):
 * row r of src [rows, ld] is copied to address dst_rows[r] unless that is 0. */
int csmoe_ep_push_rows(const void* src, int32_t dtype, int64_t ld, int32_t D, int64_t rows, const uint64_t* dst_rows,
                       void* stream);

/* Weight exchange for small experts (sigma-MoE: E*D*H*2 bytes of parameters << T*K*D of expanded token rows): the
 * parameters and their optimizer state stay sharded, every rank computes on a full operand copy.
 * gather_push: dst[r][dst_offset + i] = convert(src[i]) for every rank r (cast + all-gather in one pass; dst = peer
 * array of the full-size operand copies, dst_offset = rank * n).  f32->bf16, bf16->bf16, f32->f32.  n % 8 == 0.
 * reduce_pull: out[i] = sum over r ascending of src[r][src_offset + i] (src = peer array of full-size FP32 gradient
 * buffers, src_offset = rank * n): a deterministic reduce-scatter; out f32 or bf16.  n % 4 == 0.
 * Both need a csmoe_ep_barrier between the writers and the readers; no counterpart in the reference (data-parallel
 * all-reduce of every gradient, simple_task.py:403-413). */
int csmoe_ep_gather_push(const void* src, int32_t src_dtype, int64_t n, const void* const* dst, int32_t dst_dtype,
                         int64_t dst_offset, int32_t rank, int32_t P, void* stream);
int csmoe_ep_reduce_pull(const void* const* src, int64_t src_offset, int64_t n, void* out, int32_t out_dtype, int32_t P,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSMOE_H_ */
