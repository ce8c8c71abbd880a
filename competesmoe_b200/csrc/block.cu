// The steps either side of the MoE layer in the reference's pre-LN transformer block
// (moe_pretrain_model/layers/transformer/relative_moe_transformer.py:150-159):
//
//     src2 = self.norm2(src)                 -> csmoe_layernorm_fwd: LayerNorm (fp32 statistics) written straight in the
//                                               layer's compute dtype: the fp32 LN output and the autocast cast that
//                                               follows it never round-trip through HBM
//     src3 = self.pkm(src2, id_layer=...)
//     src  = src + self.dropout(src3)        -> csmoe_combine_residual_fwd: the gate-weighted combine of the expert rows
//                                               adds the residual and applies dropout in its epilogue (router step), or
//                                               csmoe_residual_dropout_fwd on a finished layer output (competition step)
//
// Dropout uses a counter-based Philox4x32-10 stream (seed, element index): the backward pass regenerates the mask instead
// of storing it.  (The reference draws its mask from torch's generator; the streams differ, the distribution does not.)
#include "common.h"

namespace csmoe {
namespace {

constexpr int kWarps = 8;
constexpr int kMaxK = 8;
constexpr int kColBlock = 1024;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x, hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// keep mask of the 8 consecutive elements starting at element index `e0` (a multiple of 8)
__device__ __forceinline__ void keep8(unsigned long long seed, long long e0, uint32_t threshold, bool (&keep)[8]) {
  const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const unsigned long long blk = static_cast<unsigned long long>(e0) >> 2;
  const uint4 a = philox4x32_10(make_uint4(static_cast<uint32_t>(blk), static_cast<uint32_t>(blk >> 32), 0u, 0u), key);
  const uint4 b = philox4x32_10(make_uint4(static_cast<uint32_t>(blk + 1), static_cast<uint32_t>((blk + 1) >> 32), 0u, 0u), key);
  keep[0] = a.x >= threshold; keep[1] = a.y >= threshold; keep[2] = a.z >= threshold; keep[3] = a.w >= threshold;
  keep[4] = b.x >= threshold; keep[5] = b.y >= threshold; keep[6] = b.z >= threshold; keep[7] = b.w >= threshold;
}

struct Drop {
  unsigned long long seed;
  uint32_t threshold;   // keep iff random u32 >= threshold;  threshold = p * 2^32
  float scale;          // 1 / (1 - p)
  int on;
};

// ---------------------------------------------------------------------------------------------- LayerNorm
// One warp per row; the row stays in registers between the passes (D <= 2048).
template <typename TX, typename TY>
__global__ void __launch_bounds__(kWarps * 32)
layernorm_fwd_kernel(const TX* __restrict__ x, long long T, int D, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, TY* __restrict__ y, float* __restrict__ mean,
                     float* __restrict__ rstd) {
  const int lane = threadIdx.x & 31;
  const long long t = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  if (t >= T) return;
  constexpr int kMaxV = 8;    // 2048 / 256
  float v[kMaxV][8];
  const int nv = (D + 255) / 256;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxV; ++i) {
    if (i < nv) {
      const int c = i * 256 + lane * 8;
      if (c < D) {
        load8(x + t * D + c, v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      }
    }
  }
  const float mu = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxV; ++i) {
    if (i < nv) {
      const int c = i * 256 + lane * 8;
      if (c < D) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mu;
          q += d * d;
        }
      }
    }
  }
  const float rs = rsqrtf(warp_sum(q) / D + eps);
  if (lane == 0) {
    mean[t] = mu;
    rstd[t] = rs;
  }
#pragma unroll
  for (int i = 0; i < kMaxV; ++i) {
    if (i < nv) {
      const int c = i * 256 + lane * 8;
      if (c < D) {
        float g[8], b[8], o[8];
        load8(gamma + c, g);
        load8(beta + c, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mu) * rs * g[j] + b[j];
        store8(y + t * D + c, o);
      }
    }
  }
}

// dx per row + per-CTA partial column sums of dgamma / dbeta (rows of one CTA: a contiguous chunk, fixed order).
constexpr int kLnRowsPerCta = 64;   // 8 warps x 8 rows
template <typename TX, typename TG>
__global__ void __launch_bounds__(kWarps * 32)
layernorm_bwd_kernel(const TG* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma, long long T, int D,
                     TX* __restrict__ dx, float* __restrict__ part) {
  extern __shared__ float sh[];   // [kWarps][2][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sg = sh + static_cast<long long>(warp) * 2 * D;
  float* sb = sg + D;
  for (int c = lane; c < D; c += 32) {
    sg[c] = 0.f;
    sb[c] = 0.f;
  }
  __syncwarp();
  const long long row0 = static_cast<long long>(blockIdx.x) * kLnRowsPerCta + warp * (kLnRowsPerCta / kWarps);
  for (int r = 0; r < kLnRowsPerCta / kWarps; ++r) {
    const long long t = row0 + r;
    if (t >= T) break;
    const float mu = mean[t], rs = rstd[t];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane * 8; c < D; c += 256) {
      float g[8], xv[8], gm[8];
      load8(dy + t * D + c, g);
      load8(x + t * D + c, xv);
      load8(gamma + c, gm);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (xv[j] - mu) * rs, gg = g[j] * gm[j];
        s1 += gg;
        s2 += gg * xh;
        sg[c + j] += g[j] * xh;
        sb[c + j] += g[j];
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
    for (int c = lane * 8; c < D; c += 256) {
      float g[8], xv[8], gm[8], o[8];
      load8(dy + t * D + c, g);
      load8(x + t * D + c, xv);
      load8(gamma + c, gm);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (xv[j] - mu) * rs;
        o[j] = rs * (g[j] * gm[j] - s1 - xh * s2);
      }
      store8(dx + t * D + c, o);
    }
  }
  __syncthreads();
  float* out = part + static_cast<long long>(blockIdx.x) * 2 * D;
  for (int c = threadIdx.x; c < 2 * D; c += kWarps * 32) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += sh[static_cast<long long>(w) * 2 * D + c];
    out[c] = s;
  }
}

__global__ void ln_param_grad_kernel(const float* __restrict__ part, int n_ctas, int D, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 2 * D) return;
  float s = 0.f;
  for (int i = 0; i < n_ctas; ++i) s += part[static_cast<long long>(i) * 2 * D + c];
  if (c < D)
    dgamma[c] = s;
  else
    dbeta[c - D] = s;
}

// ---------------------------------------------------------------------------------------------- residual + dropout
// out = residual + dropout(v) with v rounded to TV first (the layer's output dtype), like `src + self.dropout(src3)`.
template <typename TV, typename TR>
__device__ __forceinline__ void residual_dropout8(const float (&v)[8], const TR* res_row, TR* out_row, int c, long long e0,
                                                  const Drop& dr) {
  bool keep[8] = {true, true, true, true, true, true, true, true};
  if (dr.on) keep8(dr.seed, e0, dr.threshold, keep);
  float r[8], o[8];
  load8(res_row + c, r);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float m = round_as(v[j], static_cast<const TV*>(nullptr));
    if (dr.on) m = keep[j] ? round_as(m * dr.scale, static_cast<const TV*>(nullptr)) : 0.f;
    o[j] = r[j] + m;
  }
  store8(out_row + c, o);
}

template <typename TV, typename TR>
__global__ void __launch_bounds__(256)
residual_dropout_fwd_kernel(const TV* __restrict__ v, const TR* __restrict__ res, long long T, int D, Drop dr,
                            TR* __restrict__ out) {
  const long long n8 = T * D / 8;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n8; i += static_cast<long long>(gridDim.x) * 256) {
    float x[8];
    load8(v + i * 8, x);
    residual_dropout8<TV, TR>(x, res + i * 8, out + i * 8, 0, i * 8, dr);
  }
}

// d(layer output) = keep ? g * scale : 0, written in the layer's output dtype (g = gradient of the block output, fp32 / bf16)
template <typename TG, typename TV>
__global__ void __launch_bounds__(256)
dropout_bwd_kernel(const TG* __restrict__ g, long long n, Drop dr, TV* __restrict__ dv) {
  const long long n8 = n / 8;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n8; i += static_cast<long long>(gridDim.x) * 256) {
    bool keep[8] = {true, true, true, true, true, true, true, true};
    if (dr.on) keep8(dr.seed, i * 8, dr.threshold, keep);
    float x[8];
    load8(g + i * 8, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = dr.on ? (keep[j] ? x[j] * dr.scale : 0.f) : x[j];
    store8(dv + i * 8, x);
  }
}

// combine_fwd (permute.cu) with the block's residual + dropout in the epilogue
template <typename T, int KT, typename TR>
__global__ void __launch_bounds__(kWarps * 32)
combine_residual_kernel(const T* __restrict__ y, long long Tn, int D, int K, const int32_t* __restrict__ slot_to_row,
                        const int32_t* __restrict__ sel, const float* __restrict__ w, int flags,
                        const TR* __restrict__ res, Drop dr, TR* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int nblk = (D + kColBlock - 1) / kColBlock;
  const long long units = Tn * nblk;
  for (long long u = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5); u < units;
       u += static_cast<long long>(gridDim.x) * kWarps) {
    const long long t = u / nblk;
    const int c_lo = static_cast<int>(u % nblk) * kColBlock, c_hi = min(D, c_lo + kColBlock);
    int order[KT], key[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      order[k] = k;
      key[k] = k < K ? sel[t * K + k] : 0x7fffffff;
    }
#pragma unroll
    for (int i = 1; i < KT; ++i) {
#pragma unroll
      for (int j = i; j > 0; --j) {
        if (key[j] < key[j - 1]) {
          const int tk = key[j]; key[j] = key[j - 1]; key[j - 1] = tk;
          const int to = order[j]; order[j] = order[j - 1]; order[j - 1] = to;
        }
      }
    }
    int rows[KT];
    float ws[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) {
      rows[i] = 0;
      ws[i] = 0.f;
      if (i < K) {
        const long long s = t * K + order[i];
        rows[i] = slot_to_row[s];
        float wv = w[s];
        if (flags & 2) wv = round_as(wv, static_cast<const T*>(nullptr));
        ws[i] = wv;
      }
    }
    for (int c = c_lo + lane * 8; c < c_hi; c += 256) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      typename Raw8<T>::type raw[KT];
#pragma unroll
      for (int i = 0; i < KT; ++i)
        if (i < K) raw[i] = load8_raw(y + static_cast<long long>(rows[i]) * D + c);
#pragma unroll
      for (int i = 0; i < KT; ++i) {
        if (i < K) {
          float v[8];
          unpack8(raw[i], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[j] = fmaf(ws[i], v[j], acc[j]);
            if (flags & 1) acc[j] = round_as(acc[j], static_cast<const T*>(nullptr));
          }
        }
      }
      residual_dropout8<T, TR>(acc, res + t * D, out + t * D, c, t * D + c, dr);
    }
  }
}

inline Drop make_drop(float p, uint64_t seed) {
  Drop d;
  d.seed = seed;
  d.on = p > 0.f ? 1 : 0;
  const double thr = static_cast<double>(p) * 4294967296.0;
  d.threshold = thr >= 4294967295.0 ? 0xffffffffu : static_cast<uint32_t>(thr);
  d.scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  return d;
}

inline unsigned flat_grid(long long items, int per_block) {
  const long long blocks = (items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms() > 0 ? num_sms() : 148) * 8;
  return static_cast<unsigned>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace
}  // namespace csmoe

using namespace csmoe;

extern "C" int csmoe_layernorm_fwd(const void* x, int32_t x_dtype, int64_t T, int32_t D, const float* gamma, const float* beta,
                                   float eps, void* y, int32_t y_dtype, float* mean, float* rstd, void* stream_) {
  CSMOE_CHECK_ARG(x && gamma && beta && y && mean && rstd, "csmoe_layernorm_fwd: NULL argument");
  CSMOE_CHECK_ARG(T >= 0 && D > 0 && D % 8 == 0 && D <= 2048, "csmoe_layernorm_fwd: D must be a multiple of 8, <= 2048");
  if (T == 0) return CSMOE_OK;
  cudaStream_t stream = as_stream(stream_);
  const unsigned grid = static_cast<unsigned>((T + kWarps - 1) / kWarps);
#define LN_FWD(TX, TY)                                                                                              \
  layernorm_fwd_kernel<TX, TY><<<grid, kWarps * 32, 0, stream>>>(static_cast<const TX*>(x), T, D, gamma, beta, eps, \
                                                                 static_cast<TY*>(y), mean, rstd)
  if (x_dtype == CSMOE_F32 && y_dtype == CSMOE_BF16) {
    LN_FWD(float, __nv_bfloat16);
  } else if (x_dtype == CSMOE_F32 && y_dtype == CSMOE_F32) {
    LN_FWD(float, float);
  } else if (x_dtype == CSMOE_BF16 && y_dtype == CSMOE_BF16) {
    LN_FWD(__nv_bfloat16, __nv_bfloat16);
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_layernorm_fwd: unsupported dtype combination %d -> %d", x_dtype, y_dtype);
  }
#undef LN_FWD
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int64_t csmoe_layernorm_bwd_workspace_bytes(int64_t T, int32_t D) {
  if (T < 0 || D <= 0) return -1;
  const int64_t ctas = (T + kLnRowsPerCta - 1) / kLnRowsPerCta;
  return (ctas > 0 ? ctas : 1) * 2 * D * static_cast<int64_t>(sizeof(float));
}

extern "C" int csmoe_layernorm_bwd(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* mean,
                                   const float* rstd, const float* gamma, int64_t T, int32_t D, void* dx, float* dgamma,
                                   float* dbeta, void* workspace, void* stream_) {
  CSMOE_CHECK_ARG(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && workspace, "csmoe_layernorm_bwd: NULL argument");
  CSMOE_CHECK_ARG(T >= 0 && D > 0 && D % 8 == 0 && D <= 2048, "csmoe_layernorm_bwd: D must be a multiple of 8, <= 2048");
  cudaStream_t stream = as_stream(stream_);
  const int ctas = static_cast<int>((T + kLnRowsPerCta - 1) / kLnRowsPerCta);
  const size_t smem = static_cast<size_t>(kWarps) * 2 * D * sizeof(float);
  float* part = static_cast<float*>(workspace);
  if (T > 0) {
#define LN_BWD(TX, TG)                                                                                                   \
  do {                                                                                                                   \
    static bool configured = false;                                                                                      \
    if (!configured) {                                                                                                   \
      CSMOE_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<TX, TG>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                            kWarps * 2 * 2048 * (int)sizeof(float)));                                    \
      configured = true;                                                                                                 \
    }                                                                                                                    \
    layernorm_bwd_kernel<TX, TG><<<ctas, kWarps * 32, smem, stream>>>(static_cast<const TG*>(dy), static_cast<const TX*>(x), \
                                                                      mean, rstd, gamma, T, D, static_cast<TX*>(dx), part);  \
  } while (0)
    if (x_dtype == CSMOE_F32 && dy_dtype == CSMOE_BF16) {
      LN_BWD(float, __nv_bfloat16);
    } else if (x_dtype == CSMOE_F32 && dy_dtype == CSMOE_F32) {
      LN_BWD(float, float);
    } else if (x_dtype == CSMOE_BF16 && dy_dtype == CSMOE_BF16) {
      LN_BWD(__nv_bfloat16, __nv_bfloat16);
    } else {
      CSMOE_CHECK_ARG(false, "csmoe_layernorm_bwd: unsupported dtype combination x %d, dy %d", x_dtype, dy_dtype);
    }
#undef LN_BWD
    CSMOE_CHECK_LAUNCH();
  }
  ln_param_grad_kernel<<<(2 * D + 255) / 256, 256, 0, stream>>>(part, ctas, D, dgamma, dbeta);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_residual_dropout_fwd(const void* v, int32_t v_dtype, const void* residual, int32_t res_dtype, int64_t T,
                                          int32_t D, float p, uint64_t seed, void* out, void* stream_) {
  CSMOE_CHECK_ARG(v && residual && out && T >= 0 && D > 0 && D % 8 == 0 && p >= 0.f && p < 1.f,
                  "csmoe_residual_dropout_fwd: bad arguments");
  if (T == 0) return CSMOE_OK;
  const Drop dr = make_drop(p, seed);
  cudaStream_t stream = as_stream(stream_);
  const unsigned grid = flat_grid(T * D / 8, 256);
  // flat kernel: row pointers are the tensor bases and the column index is the flat element index
#define RD(TV, TR)                                                                                                     \
  residual_dropout_fwd_kernel<TV, TR><<<grid, 256, 0, stream>>>(static_cast<const TV*>(v), static_cast<const TR*>(residual), \
                                                                T, D, dr, static_cast<TR*>(out))
  if (v_dtype == CSMOE_BF16 && res_dtype == CSMOE_F32) {
    RD(__nv_bfloat16, float);
  } else if (v_dtype == CSMOE_BF16 && res_dtype == CSMOE_BF16) {
    RD(__nv_bfloat16, __nv_bfloat16);
  } else if (v_dtype == CSMOE_F32 && res_dtype == CSMOE_F32) {
    RD(float, float);
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_residual_dropout_fwd: unsupported dtype combination %d + %d", v_dtype, res_dtype);
  }
#undef RD
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_dropout_bwd(const void* g, int32_t g_dtype, int64_t n, float p, uint64_t seed, void* dv, int32_t v_dtype,
                                 void* stream_) {
  CSMOE_CHECK_ARG(g && dv && n >= 0 && n % 8 == 0 && p >= 0.f && p < 1.f, "csmoe_dropout_bwd: bad arguments");
  if (n == 0) return CSMOE_OK;
  const Drop dr = make_drop(p, seed);
  cudaStream_t stream = as_stream(stream_);
  const unsigned grid = flat_grid(n / 8, 256);
#define DB(TG, TV) dropout_bwd_kernel<TG, TV><<<grid, 256, 0, stream>>>(static_cast<const TG*>(g), n, dr, static_cast<TV*>(dv))
  if (g_dtype == CSMOE_F32 && v_dtype == CSMOE_BF16) {
    DB(float, __nv_bfloat16);
  } else if (g_dtype == CSMOE_BF16 && v_dtype == CSMOE_BF16) {
    DB(__nv_bfloat16, __nv_bfloat16);
  } else if (g_dtype == CSMOE_F32 && v_dtype == CSMOE_F32) {
    DB(float, float);
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_dropout_bwd: unsupported dtype combination %d -> %d", g_dtype, v_dtype);
  }
#undef DB
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_combine_residual_fwd(const void* y, int32_t dtype, int64_t T, int32_t D, int32_t K,
                                          const int32_t* slot_to_row, const int32_t* sel, const float* w, int32_t flags,
                                          const void* residual, int32_t res_dtype, float p, uint64_t seed, void* out,
                                          void* stream_) {
  CSMOE_CHECK_ARG(y && slot_to_row && sel && w && residual && out, "csmoe_combine_residual_fwd: NULL argument");
  CSMOE_CHECK_ARG(dtype == CSMOE_BF16 && D > 0 && D % 8 == 0 && K >= 1 && K <= kMaxK && p >= 0.f && p < 1.f,
                  "csmoe_combine_residual_fwd: bf16 rows, D %% 8 == 0, 1 <= K <= %d, 0 <= p < 1", kMaxK);
  if (T == 0) return CSMOE_OK;
  const Drop dr = make_drop(p, seed);
  cudaStream_t stream = as_stream(stream_);
  const long long units = T * ((D + kColBlock - 1) / kColBlock);
  const unsigned grid = flat_grid(units, kWarps);
  using T16 = __nv_bfloat16;
#define CR(KT, TR)                                                                                                      \
  combine_residual_kernel<T16, KT, TR><<<grid, kWarps * 32, 0, stream>>>(static_cast<const T16*>(y), T, D, K, slot_to_row, \
                                                                         sel, w, flags, static_cast<const TR*>(residual), \
                                                                         dr, static_cast<TR*>(out))
  if (res_dtype == CSMOE_F32) {
    if (K <= 2) CR(2, float); else if (K <= 4) CR(4, float); else CR(8, float);
  } else if (res_dtype == CSMOE_BF16) {
    if (K <= 2) CR(2, T16); else if (K <= 4) CR(4, T16); else CR(8, T16);
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_combine_residual_fwd: residual must be fp32 or bf16");
  }
#undef CR
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}
