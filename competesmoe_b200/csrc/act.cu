// Elementwise expert-FFN pieces that are not (yet) fused into a GEMM epilogue: activation forward/backward
// (ReLU / GELU / GELU-tanh / SiLU / SiLU-GLU), per-expert bias gradients and the fp32 -> bf16 parameter cast used under
// autocast.  HBM-bound, 16-byte accesses.  Replaces the eager activation calls inside the reference's expert modules
// (moe_model/model/multimodal_encoder/siglip_smoe.py:85-97, Phi3MLP; moe_pretrain_model/layers/moe/moe.py:405).
#include "common.h"

namespace csmoe {
namespace {

constexpr int kThreads = 256;
constexpr int kBiasGradSplits = 16;

template <typename T>
__global__ void __launch_bounds__(kThreads)
act_fwd_kernel(const T* __restrict__ z, long long rows, long long cols, long long ldz, int act, T* __restrict__ h,
               long long ldh, const int32_t* __restrict__ tile_expert) {
  const long long vec_per_row = cols / 8;
  const long long total = rows * vec_per_row;
  for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const long long r = i / vec_per_row, c = (i % vec_per_row) * 8;
    if (tile_expert != nullptr && __ldg(tile_expert + (r >> 7)) < 0) continue;   // row tile past the routed rows
    float v[8], o[8];
    if (act == CSMOE_ACT_SILU_GLU) {
      float u[8];
      load8(z + r * ldz + c, v);          // gate
      load8(z + r * ldz + cols + c, u);   // up
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float s = round_as(act_apply(v[j], CSMOE_ACT_SILU, sizeof(T) == 2), static_cast<const T*>(nullptr));
        o[j] = u[j] * s;
      }
    } else {
      load8(z + r * ldz + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = v[j];
      act_apply_vec<8>(o, act, sizeof(T) == 2);
    }
    store8(h + r * ldh + c, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
act_bwd_kernel(const T* __restrict__ z, const T* __restrict__ dh, long long rows, long long cols, long long ldz,
               long long ldh, int act, T* __restrict__ dz, const int32_t* __restrict__ tile_expert, int tpr_log2) {
  // grid.x x block: column vectors of a row, grid.y: rows (strided) -- no 64-bit division per vector, and the saved
  // pre-activation / incoming gradient of one row are read as whole contiguous segments
  // (`tpr_log2`: threads per row of the block, a power of two <= 256, so narrow matrices keep every thread busy)
  const int tpr = 1 << tpr_log2;
  const long long c = (static_cast<long long>(blockIdx.x) * tpr + (threadIdx.x & (tpr - 1))) * 8;
  if (c >= cols) return;
  const int rpb = kThreads >> tpr_log2;
  for (long long r = static_cast<long long>(blockIdx.y) * rpb + (threadIdx.x >> tpr_log2); r < rows;
       r += static_cast<long long>(gridDim.y) * rpb) {
    if (tile_expert != nullptr && __ldg(tile_expert + (r >> 7)) < 0) continue;
    float v[8], g[8], o[8];
    load8(dh + r * ldh + c, g);
    if (act == CSMOE_ACT_SILU_GLU) {
      float u[8], du[8];
      load8(z + r * ldz + c, v);
      load8(z + r * ldz + cols + c, u);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float s = round_as(act_apply(v[j], CSMOE_ACT_SILU, sizeof(T) == 2), static_cast<const T*>(nullptr));
        du[j] = g[j] * s;
        const float ds = round_as(g[j] * u[j], static_cast<const T*>(nullptr));
        o[j] = ds * act_grad(v[j], CSMOE_ACT_SILU, sizeof(T) == 2);
      }
      store8(dz + r * ldz + c, o);
      store8(dz + r * ldz + cols + c, du);
    } else {
      load8(z + r * ldz + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = g[j];
      act_grad_vec<8>(o, v, act, sizeof(T) == 2);
      store8(dz + r * ldz + c, o);
    }
  }
}

// dbias[e][n]: grid (ceil(n / 256), E, row splits); block = 32 column-vectors x 8 row lanes.  Each split reduces a
// contiguous share of the expert's rows; with more than one split the partial sums go to `partial`
// [splits, E, n] (fp32) and bias_grad_finish_kernel adds them in split order (deterministic, no atomics).
// FUSED: g is dh, and the kernel first forms dz = dh * act'(z) (stored to `dz`, rounded to T) and sums that: the
// activation backward and the bias gradient of the first projection in one pass over dh / z (the separate bias_grad
// pass re-read the [rows, F] gradient it had just written: 229 MB at the SigLIP shape).
template <typename T, typename OutT, bool FUSED>
__global__ void __launch_bounds__(256)
bias_grad_kernel(const T* __restrict__ g, long long ldg, int n, const int32_t* __restrict__ pad_offsets, int dense,
                 long long dense_rows, OutT* __restrict__ dbias, float* __restrict__ partial,
                 const T* __restrict__ z = nullptr, long long ldz = 0, int act = 0, T* __restrict__ dz = nullptr) {
  __shared__ float red[8][32][8];
  const int e = blockIdx.y;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cx) * 8;
  long long r0, r1;
  if (dense) {
    r0 = e * dense_rows;
    r1 = r0 + dense_rows;
  } else {
    r0 = pad_offsets[e];
    r1 = pad_offsets[e + 1];
  }
  const int splits = gridDim.z;
  if (splits > 1) {
    const long long span = ((r1 - r0 + splits - 1) / splits + 7) / 8 * 8;
    r0 += blockIdx.z * span;
    r1 = r0 + span < r1 ? r0 + span : r1;
  }
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < n) {
    for (long long r = r0 + ry; r < r1; r += 8) {
      float v[8];
      load8(g + r * ldg + col, v);
      if (FUSED) {
        float zz[8];
        load8(z + r * ldz + col, zz);
        act_grad_vec<8>(v, zz, act, sizeof(T) == 2);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = round_as(v[j], static_cast<const T*>(nullptr));
        store8(dz + r * ldz + col, v);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ry][cx][j] = acc[j];
  __syncthreads();
  if (ry == 0 && col < n) {
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) s[j] += red[y][cx][j];
    }
    if (splits > 1)
      store8(partial + (static_cast<long long>(blockIdx.z) * gridDim.y + e) * n + col, s);
    else
      store8(dbias + static_cast<long long>(e) * n + col, s);
  }
}

template <typename OutT>
__global__ void __launch_bounds__(256)
bias_grad_finish_kernel(const float* __restrict__ partial, int splits, long long en, OutT* __restrict__ dbias) {
  const long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 8;
  if (i >= en) return;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int z = 0; z < splits; ++z) {
    float v[8];
    load8(partial + z * en + i, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += v[j];
  }
  store8(dbias + i, s);
}

__global__ void __launch_bounds__(kThreads)
cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long nv = n / 8;
  for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < nv;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    float v[8];
    load8(src + i * 8, v);
    store8(dst + i * 8, v);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const long long i = nv * 8 + threadIdx.x;
    dst[i] = __float2bfloat16_rn(src[i]);
  }
}

// fp32 -> three bf16 terms with hi + mid + lo == x to 24 bits: the operands of the fp32-accurate GEMM path (six bf16
// tensor-core products accumulated in fp32 reproduce an IEEE fp32 matmul to ~2^-22 relative, where a single bf16 product
// has 2^-9).  The reference's fp32 cvmm forbids TF32 (layers/cvmm.py:395 allow_tf32=False).
__global__ void __launch_bounds__(kThreads)
split_bf16x3_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ mid,
                    __nv_bfloat16* __restrict__ lo, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const float x = src[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);
    hi[i] = h;
    mid[i] = m;
    lo[i] = __float2bfloat16_rn(r2);
  }
}

inline unsigned flat_grid(long long work_items) {
  const long long blocks = (work_items + kThreads - 1) / kThreads;
  const long long cap = static_cast<long long>(num_sms() > 0 ? num_sms() : 148) * 8;
  return static_cast<unsigned>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace
}  // namespace csmoe

using namespace csmoe;

extern "C" int csmoe_act_fwd(const void* z, int32_t dtype, int64_t rows, int64_t cols, int64_t ldz, int32_t act,
                             void* h, int64_t ldh, const int32_t* tile_expert, void* stream_) {
  CSMOE_CHECK_ARG(z && h, "csmoe_act_fwd: NULL pointer");
  CSMOE_CHECK_ARG(cols > 0 && cols % 8 == 0 && ldz % 8 == 0 && ldh % 8 == 0, "csmoe_act_fwd: cols/ld must be multiples of 8");
  CSMOE_CHECK_ARG(act >= CSMOE_ACT_NONE && act <= CSMOE_ACT_SILU_GLU, "csmoe_act_fwd: bad act %d", act);
  if (rows == 0) return CSMOE_OK;
  cudaStream_t stream = as_stream(stream_);
  const unsigned grid = flat_grid(rows * (cols / 8));
  if (dtype == CSMOE_BF16) {
    act_fwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(static_cast<const __nv_bfloat16*>(z), rows, cols, ldz,
                                                                 act, static_cast<__nv_bfloat16*>(h), ldh, tile_expert);
  } else if (dtype == CSMOE_F32) {
    act_fwd_kernel<float><<<grid, kThreads, 0, stream>>>(static_cast<const float*>(z), rows, cols, ldz, act,
                                                         static_cast<float*>(h), ldh, tile_expert);
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_act_fwd: unsupported dtype %d", dtype);
  }
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_act_bwd(const void* z, const void* dh, int32_t dtype, int64_t rows, int64_t cols, int64_t ldz,
                             int64_t ldh, int32_t act, void* dz, const int32_t* tile_expert, void* stream_) {
  CSMOE_CHECK_ARG(z && dh && dz, "csmoe_act_bwd: NULL pointer");
  CSMOE_CHECK_ARG(cols > 0 && cols % 8 == 0 && ldz % 8 == 0 && ldh % 8 == 0, "csmoe_act_bwd: cols/ld must be multiples of 8");
  CSMOE_CHECK_ARG(act >= CSMOE_ACT_NONE && act <= CSMOE_ACT_SILU_GLU, "csmoe_act_bwd: bad act %d", act);
  if (rows == 0) return CSMOE_OK;
  cudaStream_t stream = as_stream(stream_);
  int tpr_log2 = 8;
  while (tpr_log2 > 0 && (1 << (tpr_log2 - 1)) >= cols / 8) --tpr_log2;
  const int tpr = 1 << tpr_log2, rpb = kThreads >> tpr_log2;
  const unsigned gx = static_cast<unsigned>((cols / 8 + tpr - 1) / tpr);
  const long long row_blocks = (rows + rpb - 1) / rpb;
  const long long want_y = (148LL * 16 + gx - 1) / gx;     // ~16 CTAs per SM in flight
  const dim3 grid(gx, static_cast<unsigned>(row_blocks < want_y ? row_blocks : want_y));
  if (dtype == CSMOE_BF16) {
    act_bwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(static_cast<const __nv_bfloat16*>(z),
                                                                 static_cast<const __nv_bfloat16*>(dh), rows, cols, ldz,
                                                                 ldh, act, static_cast<__nv_bfloat16*>(dz), tile_expert, tpr_log2);
  } else if (dtype == CSMOE_F32) {
    act_bwd_kernel<float><<<grid, kThreads, 0, stream>>>(static_cast<const float*>(z), static_cast<const float*>(dh),
                                                         rows, cols, ldz, ldh, act, static_cast<float*>(dz), tile_expert,
                                                         tpr_log2);
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_act_bwd: unsupported dtype %d", dtype);
  }
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int64_t csmoe_bias_grad_workspace_bytes(int32_t n, int32_t num_experts) {
  if (n <= 0 || num_experts <= 0) return -1;
  return static_cast<int64_t>(kBiasGradSplits) * num_experts * n * static_cast<int64_t>(sizeof(float));
}

namespace {
// Shared launcher of csmoe_bias_grad (z == nullptr) and csmoe_act_bwd_bias (z, dz given).
int bias_grad_launch(const char* what, const void* g, int32_t dtype, int64_t ldg, int32_t n, int32_t num_experts,
                     const int32_t* pad_offsets, int32_t dense, int64_t dense_rows, void* dbias, int32_t out_dtype,
                     void* workspace, const void* z, int64_t ldz, int32_t act, void* dz, void* stream_) {
  CSMOE_CHECK_ARG(g && dbias, "%s: NULL pointer", what);
  CSMOE_CHECK_ARG(dense || pad_offsets, "%s: pad_offsets required unless dense", what);
  CSMOE_CHECK_ARG(n > 0 && n % 8 == 0 && ldg % 8 == 0 && ldz % 8 == 0, "%s: n/ld must be multiples of 8", what);
  CSMOE_CHECK_ARG(num_experts >= 1 && num_experts <= 65535, "%s: bad num_experts", what);
  cudaStream_t stream = as_stream(stream_);
  // few experts x few column blocks would leave most SMs idle: split each expert's rows when a workspace is given
  const int col_blocks = (n + 255) / 256;
  const int splits = (workspace != nullptr && col_blocks * num_experts < 4 * 148) ? kBiasGradSplits : 1;
  dim3 grid(col_blocks, num_experts, splits);
  float* partial = static_cast<float*>(workspace);
  const long long en = static_cast<long long>(num_experts) * n;
  const unsigned fin_grid = static_cast<unsigned>((en / 8 + 255) / 256);
#define LAUNCH_BG(T, O)                                                                                               \
  if (z != nullptr)                                                                                                   \
    bias_grad_kernel<T, O, true><<<grid, 256, 0, stream>>>(static_cast<const T*>(g), ldg, n, pad_offsets, dense,      \
                                                           dense_rows, static_cast<O*>(dbias), partial,              \
                                                           static_cast<const T*>(z), ldz, act, static_cast<T*>(dz)); \
  else                                                                                                                \
    bias_grad_kernel<T, O, false><<<grid, 256, 0, stream>>>(static_cast<const T*>(g), ldg, n, pad_offsets, dense,     \
                                                            dense_rows, static_cast<O*>(dbias), partial);            \
  if (splits > 1) bias_grad_finish_kernel<O><<<fin_grid, 256, 0, stream>>>(partial, splits, en, static_cast<O*>(dbias))
  if (dtype == CSMOE_BF16 && out_dtype == CSMOE_BF16) {
    LAUNCH_BG(__nv_bfloat16, __nv_bfloat16);
  } else if (dtype == CSMOE_BF16 && out_dtype == CSMOE_F32) {
    LAUNCH_BG(__nv_bfloat16, float);
  } else if (dtype == CSMOE_F32 && out_dtype == CSMOE_F32) {
    LAUNCH_BG(float, float);
  } else {
    CSMOE_CHECK_ARG(false, "%s: unsupported dtype combination %d -> %d", what, dtype, out_dtype);
  }
#undef LAUNCH_BG
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}
}  // namespace

extern "C" int csmoe_bias_grad(const void* g, int32_t dtype, int64_t ldg, int32_t n, int32_t num_experts,
                               const int32_t* pad_offsets, int32_t dense, int64_t dense_rows, void* dbias,
                               int32_t out_dtype, void* workspace, void* stream_) {
  return bias_grad_launch("csmoe_bias_grad", g, dtype, ldg, n, num_experts, pad_offsets, dense, dense_rows, dbias, out_dtype,
                          workspace, nullptr, 0, 0, nullptr, stream_);
}

extern "C" int csmoe_act_bwd_bias(const void* z, const void* dh, int32_t dtype, int64_t ldz, int64_t ldh, int32_t n,
                                  int32_t num_experts, const int32_t* pad_offsets, int32_t dense, int64_t dense_rows,
                                  int32_t act, void* dz, void* dbias, int32_t out_dtype, void* workspace, void* stream_) {
  CSMOE_CHECK_ARG(z && dz, "csmoe_act_bwd_bias: NULL pointer");
  CSMOE_CHECK_ARG(act == CSMOE_ACT_RELU || act == CSMOE_ACT_GELU || act == CSMOE_ACT_GELU_TANH || act == CSMOE_ACT_SILU,
                  "csmoe_act_bwd_bias: activation %d has no elementwise backward here", act);
  return bias_grad_launch("csmoe_act_bwd_bias", dh, dtype, ldh, n, num_experts, pad_offsets, dense, dense_rows, dbias,
                          out_dtype, workspace, z, ldz, act, dz, stream_);
}

extern "C" int csmoe_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream_) {
  CSMOE_CHECK_ARG(src && dst, "csmoe_cast_f32_bf16: NULL pointer");
  CSMOE_CHECK_ARG((reinterpret_cast<uintptr_t>(src) % 16 == 0) && (reinterpret_cast<uintptr_t>(dst) % 16 == 0),
                  "csmoe_cast_f32_bf16: pointers must be 16-byte aligned");
  if (n == 0) return CSMOE_OK;
  cast_f32_bf16_kernel<<<flat_grid(n / 8 + 1), kThreads, 0, as_stream(stream_)>>>(
      src, static_cast<__nv_bfloat16*>(dst), n);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_split_f32_bf16x3(const float* src, void* hi, void* mid, void* lo, int64_t n, void* stream_) {
  CSMOE_CHECK_ARG(src && hi && mid && lo && n >= 0, "csmoe_split_f32_bf16x3: bad arguments");
  if (n == 0) return CSMOE_OK;
  split_bf16x3_kernel<<<flat_grid(n), kThreads, 0, as_stream(stream_)>>>(
      src, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(mid), static_cast<__nv_bfloat16*>(lo), n);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}
