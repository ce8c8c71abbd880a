// Competition step: neural-response score  aff[t,e] = mean_d softplus(y_e[t,d])  and its backward.
// Replaces `torch.mean(F.softplus(out_i), dim=-1)` (moe_model/model/moe/competesmoe.py:240-243) and
// `torch.mean(F.softplus(expert_outputs), dim=-1)` (moe_pretrain_model/layers/moe/competesmoe.py:399-403).
// One warp reduces one (expert, token) row of the dense expert-output buffer y[E, t_pad, D]; HBM-bound.
#include "common.h"

namespace csmoe {
namespace {

constexpr int kWarps = 8;

// softplus(x) = max(x, 0) + log(1 + exp(-|x|)): the argument of the log is in (1, 2], where the fast intrinsics are
// accurate to ~1e-7 absolute -- far below the bf16 / fp32-mean resolution of the score -- and the kernel stays HBM bound
// instead of being bound by log1pf / expf (measured: 780 -> ~200 us on [64 x 8192 x 1024]).
__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.f) + __logf(1.f + __expf(-fabsf(x))); }
// The same value with the logarithm as a degree-7 polynomial of u = exp(-|x|) in (0, 1] (max abs error 2.1e-7, the level
// of __logf itself): one MUFU op instead of two.  The score kernel alternates the two forms element by element so that
// the MUFU pipe (16 / clk / SM) and the FMA pipe share the work: 3 MUFU + ~24 FP ops per two elements.
__device__ __forceinline__ float softplus_poly(float x) {
  const float u = __expf(-fabsf(x));
  float q = -0.00837115255f;
  q = fmaf(q, u, 0.0434938998f);
  q = fmaf(q, u, -0.106850029f);
  q = fmaf(q, u, 0.176874767f);
  q = fmaf(q, u, -0.244747742f);
  q = fmaf(q, u, 0.332719284f);
  q = fmaf(q, u, -0.499971745f);
  q = fmaf(q, u, 0.999999781f);
  return fmaf(q, u, fmaxf(x, 0.f));
}
__device__ __forceinline__ float sigmoid_sp(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

template <typename T>
__global__ void __launch_bounds__(kWarps * 32)
affinity_fwd_kernel(const T* __restrict__ y, int E, long long Tn, long long t_pad, int D, int round_bf16,
                    float* __restrict__ aff) {
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(E) * Tn;
  for (long long i = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5); i < total;
       i += static_cast<long long>(gridDim.x) * kWarps) {
    const long long e = i / Tn, t = i % Tn;
    const T* row = y + (e * t_pad + t) * D;
    float s = 0.f;
    // four 16-byte loads per lane in flight (one load at a time left the kernel latency bound at 2.5 TB/s)
    for (int c0 = lane * 8; c0 < D; c0 += 1024) {
      float v[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (c0 + i * 256 < D) load8(row + c0 + i * 256, v[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (c0 + i * 256 < D) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float sp = (j & 1) ? softplus_poly(v[i][j]) : softplus_f(v[i][j]);
            s += round_bf16 ? bf16_round(sp) : sp;
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      const float m = s / static_cast<float>(D);
      aff[t * E + e] = round_bf16 ? bf16_round(m) : m;
    }
  }
}

// Finish the score from the per-(row, 64-column group) sums the GEMM epilogue wrote.
__global__ void __launch_bounds__(256)
affinity_from_rowsum_kernel(const float* __restrict__ rowsum, int groups, int E, long long Tn, long long t_pad, int D,
                            int round_bf16, float* __restrict__ aff) {
  const long long total = static_cast<long long>(E) * Tn;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * 256) {
    const long long e = i / Tn, t = i % Tn;
    const float* r = rowsum + (e * t_pad + t) * groups;
    float s = 0.f;
    for (int g = 0; g < groups; ++g) s += r[g];
    const float m = s / static_cast<float>(D);
    aff[t * E + e] = round_bf16 ? bf16_round(m) : m;
  }
}

template <typename T>
__global__ void __launch_bounds__(kWarps * 32)
affinity_bwd_kernel(const T* __restrict__ y, const float* __restrict__ daff, int E, long long Tn, long long t_pad, int D,
                    int accumulate, T* __restrict__ dy) {
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(E) * Tn;
  const float inv_d = 1.f / static_cast<float>(D);
  for (long long i = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5); i < total;
       i += static_cast<long long>(gridDim.x) * kWarps) {
    const long long e = i / Tn, t = i % Tn;
    const long long off = (e * t_pad + t) * D;
    const float g = daff[t * E + e] * inv_d;
    for (int c = lane * 8; c < D; c += 256) {
      float v[8], o[8];
      load8(y + off + c, v);
      if (accumulate) {
        load8(dy + off + c, o);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += g * sigmoid_sp(v[j]);
      store8(dy + off + c, o);
    }
  }
}


// ---- diversity loss of the selected experts' outputs (moe_model/model/moe/competesmoe.py:180-218,
// moe_pretrain_model/layers/moe/competesmoe.py:330-372): per token, the K x K cosine-similarity matrix of its selected
// rows; loss = sum of the off-diagonal entries / (T*K*K).  One warp per token; the K rows are read once.
constexpr int kMaxK = 8;

template <typename T, int KT>
__global__ void __launch_bounds__(kWarps * 32)
diversity_fwd_kernel(const T* __restrict__ y, long long Tn, long long t_pad, int D, int K,
                     const int32_t* __restrict__ sel, float* __restrict__ inv_norm, float* __restrict__ sim,
                     float* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  for (long long t = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5); t < Tn;
       t += static_cast<long long>(gridDim.x) * kWarps) {
    const T* rows[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) rows[k] = y + (static_cast<long long>(k < K ? sel[t * K + k] : 0) * t_pad + t) * D;
    float dot[KT * (KT + 1) / 2];
#pragma unroll
    for (int i = 0; i < KT * (KT + 1) / 2; ++i) dot[i] = 0.f;
    for (int c = lane * 8; c < D; c += 256) {
      float v[KT][8];
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (k < K) load8(rows[k] + c, v[k]);
#pragma unroll
      for (int k = 0, i = 0; k < KT; ++k)
#pragma unroll
        for (int l = k; l < KT; ++l, ++i)
          if (l < K) {
#pragma unroll
            for (int j = 0; j < 8; ++j) dot[i] = fmaf(v[k][j], v[l][j], dot[i]);
          }
    }
#pragma unroll
    for (int i = 0; i < KT * (KT + 1) / 2; ++i)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot[i] += __shfl_xor_sync(0xffffffffu, dot[i], o);
    float inv[KT];
#pragma unroll
    for (int k = 0, i = 0; k < KT; ++k) {
      inv[k] = 1.f / fmaxf(sqrtf(dot[i]), 1e-12f);   // F.normalize: x / max(||x||, eps)
      i += KT - k;
    }
    float off = 0.f;
#pragma unroll
    for (int k = 0, i = 0; k < KT; ++k)
#pragma unroll
      for (int l = k; l < KT; ++l, ++i)
        if (l < K) {
          const float s = dot[i] * inv[k] * inv[l];
          if (lane == 0) {
            sim[(t * K + k) * K + l] = s;
            sim[(t * K + l) * K + k] = s;
          }
          if (l != k) off += 2.f * s;
        }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (k < K) inv_norm[t * K + k] = inv[k];
      partial[t] = off;
    }
  }
}

// Fixed-order sum of partial[0..n) * scale -> out[0] (one block: deterministic).
__global__ void __launch_bounds__(1024) sum_scale_kernel(const float* __restrict__ partial, long long n, float scale,
                                                         float* __restrict__ out) {
  __shared__ float sh[1024];
  float s = 0.f;
  for (long long i = threadIdx.x; i < n; i += 1024) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0] * scale;
}

// ---- one pass that writes the whole gradient of the dense expert outputs in a competition step:
//   dy[e,t,:] = daff[t,e] * sigmoid(y[e,t,:]) / D                                   (neural-response score)
//             + [e == sel[t,k]] * ( w[t,k] * dout[t,:]                              (gate-weighted combine)
//                                   + g_div * 2/(T*K*K) * inv_k * (sum_{l != k} n_l - n_k * sum_{l != k} sim[k,l]) )
// with n_k = y[sel_k,t,:] * inv_k.  Rows t in [T, t_pad) are zeroed (the wgrad GEMMs contract over padded rows).
// Replaces three autograd branches (affinity, gather + normalize + bmm, index_copy) and the adds that joined them.
// sigmoid through one MUFU op (tanh.approx, |rel err| ~ 5e-4: below the bf16 resolution of the gradient it scales)
__device__ __forceinline__ float sigmoid_tanh(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}

// One warp per (token, block of 1024 columns).  The warp first builds nsum[t, cols] = sum_k inv_k * y[sel_k, t, cols] in
// shared memory (diversity term only), then streams the token's E rows one after the other, each row segment read and
// written once as one contiguous 2 KiB piece (the earlier column-outer order touched 512 B per row visit).  87 % of the
// rows at E = 64, K = 8 only need ga * sigmoid(y): one MUFU and a few FP ops per element, so the pass stays HBM bound.
constexpr int kColBlock = 1024;

template <typename T, int KT, bool FAST>
__global__ void __launch_bounds__(kWarps * 32)
compete_bwd_kernel(const T* __restrict__ y, int E, long long Tn, long long t_pad, int D, int K,
                   const float* __restrict__ daff, const int32_t* __restrict__ sel, const float* __restrict__ w,
                   const T* __restrict__ dout, const float* __restrict__ inv_norm, const float* __restrict__ sim,
                   const float* __restrict__ g_div, T* __restrict__ dy) {
  __shared__ float nsum_all[kWarps][kColBlock];        // diversity term only
  const int lane = threadIdx.x & 31;
  float* nsum = nsum_all[threadIdx.x >> 5];
  const float inv_d = 1.f / static_cast<float>(D);
  const float gd = g_div ? g_div[0] * 2.f / (static_cast<float>(Tn) * K * K) : 0.f;
  const int nblk = (D + kColBlock - 1) / kColBlock;
  const long long units = t_pad * nblk;
  for (long long u = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5); u < units;
       u += static_cast<long long>(gridDim.x) * kWarps) {
    const long long t = u / nblk;
    const int c_lo = static_cast<int>(u % nblk) * kColBlock;
    const int c_hi = min(D, c_lo + kColBlock);
    if (t >= Tn) {
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int e = 0; e < E; ++e)
        for (int c = c_lo + lane * 8; c < c_hi; c += 256) store8(dy + (e * t_pad + t) * D + c, z);
      continue;
    }
    int se[KT];
    float wk[KT], ik[KT], ak[KT], rs[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      se[k] = k < K ? sel[t * K + k] : -1;
      wk[k] = (k < K && dout) ? w[t * K + k] : 0.f;
      ik[k] = (k < K && g_div) ? inv_norm[t * K + k] : 0.f;
      ak[k] = gd * ik[k];
      rs[k] = 0.f;
      if (k < K && g_div)
        for (int l = 0; l < K; ++l)
          if (l != k) rs[k] += sim[(t * K + k) * K + l];
    }
    if (g_div) {
      __syncwarp();
      for (int c = c_lo + lane * 8; c < c_hi; c += 256) {
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < KT; ++k)
          if (k < K) {
            float v[8];
            load8(y + (static_cast<long long>(se[k]) * t_pad + t) * D + c, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fmaf(v[j], ik[k], a[j]);
          }
        *reinterpret_cast<float4*>(nsum + (c - c_lo)) = make_float4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<float4*>(nsum + (c - c_lo) + 4) = make_float4(a[4], a[5], a[6], a[7]);
      }
      __syncwarp();
    }
#pragma unroll 1
    for (int e = 0; e < E; ++e) {
      const long long off = (e * t_pad + t) * D;
      const float ga = daff ? daff[t * E + e] * inv_d : 0.f;
      float wsel = 0.f, asel = 0.f, isel = 0.f, rsel = 0.f;
      bool picked = false;
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (se[k] == e) { wsel = wk[k]; asel = ak[k]; isel = ik[k]; rsel = rs[k]; picked = true; }
      if (!picked) {
#pragma unroll 4
        for (int c = c_lo + lane * 8; c < c_hi; c += 256) {
          float v[8];
          load8(y + off + c, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = ga * (FAST ? sigmoid_tanh(v[j]) : sigmoid_sp(v[j]));
          store8(dy + off + c, v);
        }
      } else {
#pragma unroll 2
        for (int c = c_lo + lane * 8; c < c_hi; c += 256) {
          float v[8], o[8], g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ns[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          load8(y + off + c, v);
          if (dout) load8(dout + t * D + c, g);
          if (g_div) {
            const float4 n0 = *reinterpret_cast<const float4*>(nsum + (c - c_lo));
            const float4 n1 = *reinterpret_cast<const float4*>(nsum + (c - c_lo) + 4);
            ns[0] = n0.x, ns[1] = n0.y, ns[2] = n0.z, ns[3] = n0.w, ns[4] = n1.x, ns[5] = n1.y, ns[6] = n1.z, ns[7] = n1.w;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float n = v[j] * isel;
            o[j] = ga * (FAST ? sigmoid_tanh(v[j]) : sigmoid_sp(v[j])) + wsel * g[j] + asel * (ns[j] - n - n * rsel);
          }
          store8(dy + off + c, o);
        }
      }
    }
  }
}

inline unsigned row_grid(long long rows) {
  const long long blocks = (rows + kWarps - 1) / kWarps;
  const long long cap = static_cast<long long>(num_sms() > 0 ? num_sms() : 148) * 16;
  return static_cast<unsigned>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace
}  // namespace csmoe

using namespace csmoe;

extern "C" int csmoe_affinity_fwd(const void* y, int32_t dtype, int32_t E, int64_t T_, int64_t t_pad, int32_t D,
                                  int32_t round_dtype, float* aff, void* stream_) {
  CSMOE_CHECK_ARG(y && aff, "csmoe_affinity_fwd: NULL pointer");
  CSMOE_CHECK_ARG(E >= 1 && D > 0 && D % 8 == 0 && t_pad >= T_, "csmoe_affinity_fwd: bad sizes");
  if (T_ == 0) return CSMOE_OK;
  cudaStream_t stream = as_stream(stream_);
  const unsigned grid = row_grid(static_cast<long long>(E) * T_);
  const int rb = round_dtype == CSMOE_BF16 ? 1 : 0;
  if (dtype == CSMOE_BF16) {
    affinity_fwd_kernel<__nv_bfloat16><<<grid, kWarps * 32, 0, stream>>>(static_cast<const __nv_bfloat16*>(y), E, T_,
                                                                         t_pad, D, rb, aff);
  } else if (dtype == CSMOE_F32) {
    affinity_fwd_kernel<float><<<grid, kWarps * 32, 0, stream>>>(static_cast<const float*>(y), E, T_, t_pad, D, rb, aff);
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_affinity_fwd: unsupported dtype %d", dtype);
  }
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_affinity_from_rowsum(const float* rowsum, int32_t groups, int32_t E, int64_t T_, int64_t t_pad,
                                          int32_t D, int32_t round_dtype, float* aff, void* stream_) {
  CSMOE_CHECK_ARG(rowsum && aff, "csmoe_affinity_from_rowsum: NULL pointer");
  CSMOE_CHECK_ARG(E >= 1 && D > 0 && groups == (D + 63) / 64 && t_pad >= T_, "csmoe_affinity_from_rowsum: bad sizes");
  if (T_ == 0) return CSMOE_OK;
  const long long total = static_cast<long long>(E) * T_;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  affinity_from_rowsum_kernel<<<grid, 256, 0, as_stream(stream_)>>>(rowsum, groups, E, T_, t_pad, D,
                                                                    round_dtype == CSMOE_BF16 ? 1 : 0, aff);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_affinity_bwd(const void* y, const float* daff, int32_t dtype, int32_t E, int64_t T_, int64_t t_pad,
                                  int32_t D, int32_t accumulate, void* dy, void* stream_) {
  CSMOE_CHECK_ARG(y && daff && dy, "csmoe_affinity_bwd: NULL pointer");
  CSMOE_CHECK_ARG(E >= 1 && D > 0 && D % 8 == 0 && t_pad >= T_, "csmoe_affinity_bwd: bad sizes");
  if (T_ == 0) return CSMOE_OK;
  cudaStream_t stream = as_stream(stream_);
  const unsigned grid = row_grid(static_cast<long long>(E) * T_);
  if (dtype == CSMOE_BF16) {
    affinity_bwd_kernel<__nv_bfloat16><<<grid, kWarps * 32, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(y), daff, E, T_, t_pad, D, accumulate, static_cast<__nv_bfloat16*>(dy));
  } else if (dtype == CSMOE_F32) {
    affinity_bwd_kernel<float><<<grid, kWarps * 32, 0, stream>>>(static_cast<const float*>(y), daff, E, T_, t_pad, D,
                                                                 accumulate, static_cast<float*>(dy));
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_affinity_bwd: unsupported dtype %d", dtype);
  }
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

#define DISPATCH_DTYPE_K(dtype, K, ...)                                      \
  if ((dtype) == CSMOE_BF16) {                                               \
    using T = __nv_bfloat16;                                                 \
    if ((K) <= 2) { constexpr int KT = 2; __VA_ARGS__; }                     \
    else if ((K) <= 4) { constexpr int KT = 4; __VA_ARGS__; }                \
    else { constexpr int KT = 8; __VA_ARGS__; }                              \
  } else if ((dtype) == CSMOE_F32) {                                         \
    using T = float;                                                         \
    if ((K) <= 2) { constexpr int KT = 2; __VA_ARGS__; }                     \
    else if ((K) <= 4) { constexpr int KT = 4; __VA_ARGS__; }                \
    else { constexpr int KT = 8; __VA_ARGS__; }                              \
  } else {                                                                   \
    CSMOE_CHECK_ARG(false, "unsupported dtype %d", (dtype));                 \
  }

#define DISPATCH_K(K, ...)                                 \
  if ((K) <= 2) { constexpr int KT = 2; __VA_ARGS__; }     \
  else if ((K) <= 4) { constexpr int KT = 4; __VA_ARGS__; } \
  else { constexpr int KT = 8; __VA_ARGS__; }

extern "C" int csmoe_diversity_fwd(const void* y, int32_t dtype, int64_t T_, int64_t t_pad, int32_t D, int32_t K,
                                   const int32_t* sel, float* inv_norm, float* sim, float* partial, float* loss,
                                   void* stream_) {
  CSMOE_CHECK_ARG(y && sel && inv_norm && sim && partial && loss, "csmoe_diversity_fwd: NULL pointer");
  CSMOE_CHECK_ARG(D > 0 && D % 8 == 0 && K >= 1 && K <= kMaxK && t_pad >= T_,
                  "csmoe_diversity_fwd: D %% 8 == 0, 1 <= K <= %d, t_pad >= T", kMaxK);
  cudaStream_t stream = as_stream(stream_);
  if (T_ > 0) {
    DISPATCH_DTYPE_K(dtype, K, (diversity_fwd_kernel<T, KT><<<row_grid(T_), kWarps * 32, 0, stream>>>(
                                   static_cast<const T*>(y), T_, t_pad, D, K, sel, inv_norm, sim, partial)));
    CSMOE_CHECK_LAUNCH();
  }
  const float scale = T_ > 0 ? 1.f / (static_cast<float>(T_) * K * K) : 0.f;
  sum_scale_kernel<<<1, 1024, 0, stream>>>(partial, T_, scale, loss);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_compete_bwd(const void* y, int32_t dtype, int32_t E, int64_t T_, int64_t t_pad, int32_t D,
                                 int32_t K, const float* daff, const int32_t* sel, const float* w, const void* dout,
                                 const float* inv_norm, const float* sim, const float* g_div, void* dy, void* stream_) {
  CSMOE_CHECK_ARG(y && sel && dy, "csmoe_compete_bwd: NULL pointer");
  CSMOE_CHECK_ARG(!dout || w, "csmoe_compete_bwd: dout needs the combine weights");
  CSMOE_CHECK_ARG(!g_div || (inv_norm && sim), "csmoe_compete_bwd: g_div needs the saved norms and similarities");
  CSMOE_CHECK_ARG(E >= 1 && D > 0 && D % 8 == 0 && K >= 1 && K <= kMaxK && t_pad >= T_,
                  "csmoe_compete_bwd: D %% 8 == 0, 1 <= K <= %d, t_pad >= T", kMaxK);
  if (t_pad == 0) return CSMOE_OK;
  cudaStream_t stream = as_stream(stream_);
  // bf16 gradients take the one-MUFU sigmoid; fp32 ones keep the exp / divide form
  const unsigned grid = row_grid(t_pad * ((D + kColBlock - 1) / kColBlock));
  if (dtype == CSMOE_BF16) {
    DISPATCH_K(K, (compete_bwd_kernel<__nv_bfloat16, KT, true><<<grid, kWarps * 32, 0, stream>>>(
                      static_cast<const __nv_bfloat16*>(y), E, T_, t_pad, D, K, daff, sel, w,
                      static_cast<const __nv_bfloat16*>(dout), inv_norm, sim, g_div, static_cast<__nv_bfloat16*>(dy))));
  } else if (dtype == CSMOE_F32) {
    DISPATCH_K(K, (compete_bwd_kernel<float, KT, false><<<grid, kWarps * 32, 0, stream>>>(
                      static_cast<const float*>(y), E, T_, t_pad, D, K, daff, sel, w, static_cast<const float*>(dout),
                      inv_norm, sim, g_div, static_cast<float*>(dy))));
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_compete_bwd: unsupported dtype %d", dtype);
  }
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}
