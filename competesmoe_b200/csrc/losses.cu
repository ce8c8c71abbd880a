// Competition-step losses and the small [T, E] / [T, K] backward pieces that used to run as ~40 eager PyTorch launches:
//
//   csmoe_losses_fwd / bwd     softmax of the affinity scores, the router-distillation MSE in all its variants (plain,
//                              in_topk, hybrid, tribrid), the multimodal balance loss on the affinity and the pretrain
//                              entropy balance on the affinity softmax -- one pass over [T, E] + a fixed-order reduction.
//                              moe_model/model/moe/competesmoe.py:322-335,350-371 and moe.py:90-110;
//                              moe_pretrain_model/layers/moe/competesmoe.py:541-593 and layers/moe/moe.py:323-332.
//   csmoe_entropy_balance_*    router-step regulariser of the pretrain layer, `entropy_balance(gate_logits)`
//                              (layers/moe/moe.py:323-332), from the probabilities the router kernel already produced.
//   csmoe_topk_renorm_bwd      backward of w = v / sum(v), v = scores (or sigmoid(scores)) at the top-k indices
//                              (competesmoe.py:249-254: the weights are not detached).
//   csmoe_dense_rows           row index of every selected (token, expert) pair inside the dense [E, t_pad, D] outputs.
//
// Every reduction is two-stage with a fixed summation order (no atomics): results are bit-reproducible run to run.
#include "common.h"

namespace csmoe {
namespace {

constexpr int kWarps = 8;
constexpr int kTokPerWarp = 16;
constexpr int kTokPerBlock = kWarps * kTokPerWarp;   // 128 tokens of one batch row per CTA
constexpr int kMaxE = 64;                            // two experts per lane, like the router kernels
constexpr int kMaxK = 8;
constexpr int kPart = 3 * kMaxE + 4;                 // per-CTA partials: colq[E] cnt[E] colr[E] s0 s1 s2 (+pad)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// softmax over the (up to 64) values a warp holds two per lane
__device__ __forceinline__ void warp_softmax(float a0, float a1, bool h0, bool h1, float& q0, float& q1) {
  const float m = warp_max(fmaxf(h0 ? a0 : -INFINITY, h1 ? a1 : -INFINITY));
  const float e0 = h0 ? expf(a0 - m) : 0.f, e1 = h1 ? expf(a1 - m) : 0.f;
  const float d = warp_sum(e0 + e1);
  q0 = e0 / d;
  q1 = e1 / d;
}

__device__ __forceinline__ bool in_list(const int32_t* idx, int K, int e) {
  bool hit = false;
  for (int k = 0; k < K; ++k) hit |= (idx[k] == e);
  return hit;
}

// ---- stage 1: grid (chunks, B).  Warp w of CTA c walks tokens c*128 + w*16 + i of batch row b in order.
// entropy_only: `p` holds probabilities; only their column sums are produced (colr), nothing else is read.
template <bool kEntropyOnly>
__global__ void __launch_bounds__(kWarps * 32)
losses_stage1(const float* __restrict__ p, const float* __restrict__ aff, const int32_t* __restrict__ aff_idx,
              const int32_t* __restrict__ gate_idx, long long N, int E, int K, float* __restrict__ q,
              float* __restrict__ part) {
  __shared__ float sh[kWarps][kPart];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.y;
  const bool h0 = lane < E, h1 = lane + 32 < E;
  float cq0 = 0.f, cq1 = 0.f, cn0 = 0.f, cn1 = 0.f, cr0 = 0.f, cr1 = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
  if (kEntropyOnly) {
    // column sums only: all sixteen rows' loads are issued before the first add (the rolled loop was a chain of
    // dependent ~0.6 us loads: 10.8 us for 8192 x 64 probabilities)
    float v0[kTokPerWarp], v1[kTokPerWarp];
#pragma unroll
    for (int i = 0; i < kTokPerWarp; ++i) {
      const long long n = static_cast<long long>(blockIdx.x) * kTokPerBlock + warp * kTokPerWarp + i;
      const long long t = b * N + n;
      v0[i] = (n < N && h0) ? p[t * E + lane] : 0.f;
      v1[i] = (n < N && h1) ? p[t * E + lane + 32] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kTokPerWarp; ++i) {
      cr0 += v0[i];
      cr1 += v1[i];
    }
  }
  for (int i = 0; i < (kEntropyOnly ? 0 : kTokPerWarp); ++i) {
    const long long n = static_cast<long long>(blockIdx.x) * kTokPerBlock + warp * kTokPerWarp + i;
    if (n >= N) break;
    const long long t = b * N + n;
    const float p0 = h0 ? p[t * E + lane] : 0.f, p1 = h1 ? p[t * E + lane + 32] : 0.f;
    const float a0 = h0 ? aff[t * E + lane] : 0.f, a1 = h1 ? aff[t * E + lane + 32] : 0.f;
    float q0, q1, r0, r1;
    warp_softmax(a0, a1, h0, h1, q0, q1);
    warp_softmax(q0, q1, h0, h1, r0, r1);       // entropy_balance applies log_softmax to the affinity SOFTMAX (:542-545)
    if (h0) q[t * E + lane] = q0;
    if (h1) q[t * E + lane + 32] = q1;
    const float d0 = p0 - q0, d1 = p1 - q1;
    const float sq0 = h0 ? d0 * d0 : 0.f, sq1 = h1 ? d1 * d1 : 0.f;
    s0 += sq0 + sq1;
    const int32_t* ai = aff_idx + t * K;
    s1 += (in_list(ai, K, lane) ? sq0 : 0.f) + (in_list(ai, K, lane + 32) ? sq1 : 0.f);
    if (gate_idx != nullptr) {
      const int32_t* gi = gate_idx + t * K;
      s2 += (in_list(gi, K, lane) ? sq0 : 0.f) + (in_list(gi, K, lane + 32) ? sq1 : 0.f);
    }
    const int top1 = ai[0];
    cq0 += q0; cq1 += q1;
    cr0 += r0; cr1 += r1;
    cn0 += (top1 == lane) ? 1.f : 0.f;
    cn1 += (top1 == lane + 32) ? 1.f : 0.f;
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
  float* mine = sh[warp];
  mine[lane] = cq0; mine[lane + 32] = cq1;
  mine[kMaxE + lane] = cn0; mine[kMaxE + lane + 32] = cn1;
  mine[2 * kMaxE + lane] = cr0; mine[2 * kMaxE + lane + 32] = cr1;
  if (lane == 0) { mine[3 * kMaxE] = s0; mine[3 * kMaxE + 1] = s1; mine[3 * kMaxE + 2] = s2; mine[3 * kMaxE + 3] = 0.f; }
  __syncthreads();
  float* out = part + (b * gridDim.x + blockIdx.x) * kPart;
  for (int j = threadIdx.x; j < kPart; j += kWarps * 32) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += sh[w][j];
    out[j] = s;
  }
}

// ---- stage 2: ONE CTA.  Sums the per-CTA partials of every batch row in chunk order, derives the [B, E] column
// statistics and the five scalars with a fixed-shape tree.
__device__ float block_sum_1024(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = (threadIdx.x < 32) ? red[threadIdx.x] : 0.f;
  if (warp == 0) r = warp_sum(r);
  if (threadIdx.x == 0) red[32] = r;
  __syncthreads();
  return red[32];
}

template <bool kEntropyOnly>
__global__ void __launch_bounds__(1024)
losses_stage2(const float* __restrict__ part, int chunks, long long B, long long N, int E, int K,
              float* __restrict__ colq, float* __restrict__ cnt, float* __restrict__ colr, float* __restrict__ losses) {
  __shared__ float red[33];
  float bal = 0.f, ent = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
  const float invN = 1.f / static_cast<float>(N);
  for (long long i = threadIdx.x; i < B * E; i += blockDim.x) {
    const long long b = i / E;
    const int e = static_cast<int>(i % E);
    float a = 0.f, c = 0.f, r = 0.f;
    for (int ch = 0; ch < chunks; ++ch) {
      const float* pp = part + (b * chunks + ch) * kPart;
      if (!kEntropyOnly) { a += pp[e]; c += pp[kMaxE + e]; }
      r += pp[2 * kMaxE + e];
    }
    if (!kEntropyOnly) { colq[i] = a; cnt[i] = c; bal += (a * invN) * (c * invN); }
    colr[i] = r;
    const float m = r * invN;
    ent += m > 0.f ? m * logf(m) : 0.f;
  }
  if (!kEntropyOnly) {
    for (long long i = threadIdx.x; i < B * chunks; i += blockDim.x) {
      const float* pp = part + i * kPart + 3 * kMaxE;
      s0 += pp[0]; s1 += pp[1]; s2 += pp[2];
    }
  }
  bal = block_sum_1024(bal, red);
  ent = block_sum_1024(ent, red);
  if (!kEntropyOnly) {
    s0 = block_sum_1024(s0, red);
    s1 = block_sum_1024(s1, red);
    s2 = block_sum_1024(s2, red);
  }
  if (threadIdx.x == 0) {
    const float T = static_cast<float>(B) * static_cast<float>(N);
    if (kEntropyOnly) {
      losses[0] = ent / static_cast<float>(B);
    } else {
      losses[0] = s0 / (T * E);
      losses[1] = s1 / (T * K);
      losses[2] = s2 / (T * K);
      losses[3] = bal / (static_cast<float>(B) * E) * static_cast<float>(E) * static_cast<float>(E);
      losses[4] = ent / static_cast<float>(B);
    }
  }
}

// ---- backward: one warp per token.
__global__ void __launch_bounds__(kWarps * 32)
losses_bwd_kernel(const float* __restrict__ p, const float* __restrict__ q, const int32_t* __restrict__ aff_idx,
                  const int32_t* __restrict__ gate_idx, const float* __restrict__ cnt, const float* __restrict__ colr,
                  const float* __restrict__ g, long long B, long long N, int E, int K, float* __restrict__ dp,
                  float* __restrict__ daff) {
  const int lane = threadIdx.x & 31;
  const long long t = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  const long long T = B * N;
  if (t >= T) return;
  const long long b = t / N;
  const bool h0 = lane < E, h1 = lane + 32 < E;
  const float g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3], g4 = g[4];
  const float p0 = h0 ? p[t * E + lane] : 0.f, p1 = h1 ? p[t * E + lane + 32] : 0.f;
  const float q0 = h0 ? q[t * E + lane] : 0.f, q1 = h1 ? q[t * E + lane + 32] : 0.f;
  const float Tf = static_cast<float>(T), Nf = static_cast<float>(N), Bf = static_cast<float>(B);
  const int32_t* ai = aff_idx + t * K;
  float c0 = g0 / (Tf * E), c1 = c0;
  if (in_list(ai, K, lane)) c0 += g1 / (Tf * K);
  if (in_list(ai, K, lane + 32)) c1 += g1 / (Tf * K);
  if (gate_idx != nullptr) {
    const int32_t* gi = gate_idx + t * K;
    if (in_list(gi, K, lane)) c0 += g2 / (Tf * K);
    if (in_list(gi, K, lane + 32)) c1 += g2 / (Tf * K);
  }
  if (h0) dp[t * E + lane] = 2.f * (p0 - q0) * c0;
  if (h1) dp[t * E + lane + 32] = 2.f * (p1 - q1) * c1;
  // through q: balance (linear in q) + entropy of the batch-mean of softmax(q)
  const float kb = g3 * static_cast<float>(E) / (Bf * Nf * Nf);
  float dq0 = h0 ? kb * cnt[b * E + lane] : 0.f, dq1 = h1 ? kb * cnt[b * E + lane + 32] : 0.f;
  if (g4 != 0.f) {
    float r0, r1;
    warp_softmax(q0, q1, h0, h1, r0, r1);
    const float m0 = h0 ? colr[b * E + lane] / Nf : 1.f, m1 = h1 ? colr[b * E + lane + 32] / Nf : 1.f;
    const float dr0 = h0 && m0 > 0.f ? g4 * (logf(m0) + 1.f) / (Bf * Nf) : 0.f;
    const float dr1 = h1 && m1 > 0.f ? g4 * (logf(m1) + 1.f) / (Bf * Nf) : 0.f;
    const float dot = warp_sum(r0 * dr0 + r1 * dr1);
    dq0 += r0 * (dr0 - dot);
    dq1 += r1 * (dr1 - dot);
  }
  const float dotq = warp_sum(q0 * dq0 + q1 * dq1);
  if (h0) daff[t * E + lane] = q0 * (dq0 - dotq);
  if (h1) daff[t * E + lane + 32] = q1 * (dq1 - dotq);
}

// dprobs[t, e] = g * (log m[b, e] + 1) / (B * N)
__global__ void entropy_balance_bwd_kernel(const float* __restrict__ colr, const float* __restrict__ g, long long B,
                                           long long N, int E, float* __restrict__ dprobs) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= B * N * E) return;
  const long long b = i / (N * E);
  const int e = static_cast<int>(i % E);
  const float m = colr[b * E + e] / static_cast<float>(N);
  dprobs[i] = m > 0.f ? g[0] * (logf(m) + 1.f) / (static_cast<float>(B) * static_cast<float>(N)) : 0.f;
}

// one thread per token
__global__ void topk_renorm_bwd_kernel(const float* __restrict__ scores, const float* __restrict__ w,
                                       const int32_t* __restrict__ idx, const float* __restrict__ dw, long long T, int E,
                                       int K, int sigmoid, int accumulate, float* __restrict__ dscores) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float v[kMaxK], dv[kMaxK];
  float s = 0.f, dot = 0.f;
  for (int k = 0; k < K; ++k) {
    float x = scores[t * E + idx[t * K + k]];
    if (sigmoid) x = 1.f / (1.f + expf(-x));
    v[k] = x;
    s += x;
    dot += dw[t * K + k] * w[t * K + k];
  }
  for (int k = 0; k < K; ++k) {
    float d = (dw[t * K + k] - dot) / s;
    if (sigmoid) d *= v[k] * (1.f - v[k]);
    dv[k] = d;
  }
  if (!accumulate)
    for (int e = 0; e < E; ++e) dscores[t * E + e] = 0.f;
  for (int k = 0; k < K; ++k) dscores[t * E + idx[t * K + k]] += dv[k];
}

__global__ void dense_rows_kernel(const int32_t* __restrict__ idx, long long T, int K, long long t_pad,
                                  int32_t* __restrict__ rows) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= T * K) return;
  rows[i] = static_cast<int32_t>(static_cast<long long>(idx[i]) * t_pad + i / K);
}

// ------------------------------------------------------------------------------------------------ more than 64 experts
// 64 < E <= 256 (the pretrain plugin's default -moe.n_experts is 128): the kernels above with C = 4 or 8 experts per lane
// (lane l: experts l, l + 32, ...), partials [3 * 32 C + 4] per CTA.  Same formulas, same fixed-order reductions.
constexpr int kMaxEWide = 256;
__host__ __device__ constexpr int part_stride(int C) { return 3 * 32 * C + 4; }
inline int wide_slots(int E) { return E <= 128 ? 4 : 8; }

template <int C>
__device__ __forceinline__ void warp_softmax_wide(const float (&a)[C], int lane, int E, float (&q)[C]) {
  float m = -INFINITY;
#pragma unroll
  for (int c = 0; c < C; ++c)
    if (lane + 32 * c < E) m = fmaxf(m, a[c]);
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    q[c] = lane + 32 * c < E ? expf(a[c] - m) : 0.f;
    s += q[c];
  }
  const float d = warp_sum(s);
#pragma unroll
  for (int c = 0; c < C; ++c) q[c] = q[c] / d;
}

template <bool kEntropyOnly, int C>
__global__ void __launch_bounds__(kWarps * 32)
losses_stage1_wide(const float* __restrict__ p, const float* __restrict__ aff, const int32_t* __restrict__ aff_idx,
                   const int32_t* __restrict__ gate_idx, long long N, int E, int K, float* __restrict__ q,
                   float* __restrict__ part) {
  constexpr int kP = part_stride(C), kME = 32 * C;
  __shared__ float sh[kWarps][kP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.y;
  float cq[C], cn[C], cr[C], s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) cq[c] = cn[c] = cr[c] = 0.f;
  for (int i = 0; i < kTokPerWarp; ++i) {
    const long long n = static_cast<long long>(blockIdx.x) * kTokPerBlock + warp * kTokPerWarp + i;
    if (n >= N) break;
    const long long t = b * N + n;
    if constexpr (kEntropyOnly) {
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (lane + 32 * c < E) cr[c] += p[t * E + lane + 32 * c];
    } else {
    float pv[C], av[C], qv[C], rv[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const bool h = lane + 32 * c < E;
      pv[c] = h ? p[t * E + lane + 32 * c] : 0.f;
      av[c] = h ? aff[t * E + lane + 32 * c] : 0.f;
    }
    warp_softmax_wide<C>(av, lane, E, qv);
    warp_softmax_wide<C>(qv, lane, E, rv);      // entropy_balance applies log_softmax to the affinity SOFTMAX (:542-545)
    const int32_t* ai = aff_idx + t * K;
    const int32_t* gi = gate_idx != nullptr ? gate_idx + t * K : nullptr;
    const int top1 = ai[0];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int e = lane + 32 * c;
      const bool h = e < E;
      if (h) q[t * E + e] = qv[c];
      const float d = pv[c] - qv[c];
      const float sq = h ? d * d : 0.f;
      s0 += sq;
      s1 += in_list(ai, K, e) ? sq : 0.f;
      if (gi != nullptr) s2 += in_list(gi, K, e) ? sq : 0.f;
      cq[c] += qv[c];
      cr[c] += rv[c];
      cn[c] += (top1 == e) ? 1.f : 0.f;
    }
    }
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
  float* mine = sh[warp];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    mine[lane + 32 * c] = cq[c];
    mine[kME + lane + 32 * c] = cn[c];
    mine[2 * kME + lane + 32 * c] = cr[c];
  }
  if (lane == 0) { mine[3 * kME] = s0; mine[3 * kME + 1] = s1; mine[3 * kME + 2] = s2; mine[3 * kME + 3] = 0.f; }
  __syncthreads();
  float* out = part + (b * gridDim.x + blockIdx.x) * kP;
  for (int j = threadIdx.x; j < kP; j += kWarps * 32) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += sh[w][j];
    out[j] = s;
  }
}

// stage 2 with the partial stride as an argument (me = 32 C experts per section)
template <bool kEntropyOnly>
__global__ void __launch_bounds__(1024)
losses_stage2_wide(const float* __restrict__ part, int me, int chunks, long long B, long long N, int E, int K,
                   float* __restrict__ colq, float* __restrict__ cnt, float* __restrict__ colr, float* __restrict__ losses) {
  __shared__ float red[33];
  const int kP = 3 * me + 4;
  float bal = 0.f, ent = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
  const float invN = 1.f / static_cast<float>(N);
  for (long long i = threadIdx.x; i < B * E; i += blockDim.x) {
    const long long b = i / E;
    const int e = static_cast<int>(i % E);
    float a = 0.f, c = 0.f, r = 0.f;
    for (int ch = 0; ch < chunks; ++ch) {
      const float* pp = part + (b * chunks + ch) * kP;
      if (!kEntropyOnly) { a += pp[e]; c += pp[me + e]; }
      r += pp[2 * me + e];
    }
    if (!kEntropyOnly) { colq[i] = a; cnt[i] = c; bal += (a * invN) * (c * invN); }
    colr[i] = r;
    const float m = r * invN;
    ent += m > 0.f ? m * logf(m) : 0.f;
  }
  if (!kEntropyOnly) {
    for (long long i = threadIdx.x; i < B * chunks; i += blockDim.x) {
      const float* pp = part + i * kP + 3 * me;
      s0 += pp[0]; s1 += pp[1]; s2 += pp[2];
    }
  }
  bal = block_sum_1024(bal, red);
  ent = block_sum_1024(ent, red);
  if (!kEntropyOnly) {
    s0 = block_sum_1024(s0, red);
    s1 = block_sum_1024(s1, red);
    s2 = block_sum_1024(s2, red);
  }
  if (threadIdx.x == 0) {
    const float T = static_cast<float>(B) * static_cast<float>(N);
    if (kEntropyOnly) {
      losses[0] = ent / static_cast<float>(B);
    } else {
      losses[0] = s0 / (T * E);
      losses[1] = s1 / (T * K);
      losses[2] = s2 / (T * K);
      losses[3] = bal / (static_cast<float>(B) * E) * static_cast<float>(E) * static_cast<float>(E);
      losses[4] = ent / static_cast<float>(B);
    }
  }
}

template <int C>
__global__ void __launch_bounds__(kWarps * 32)
losses_bwd_wide_kernel(const float* __restrict__ p, const float* __restrict__ q, const int32_t* __restrict__ aff_idx,
                       const int32_t* __restrict__ gate_idx, const float* __restrict__ cnt, const float* __restrict__ colr,
                       const float* __restrict__ g, long long B, long long N, int E, int K, float* __restrict__ dp,
                       float* __restrict__ daff) {
  const int lane = threadIdx.x & 31;
  const long long t = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  const long long T = B * N;
  if (t >= T) return;
  const long long b = t / N;
  const float g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3], g4 = g[4];
  const float Tf = static_cast<float>(T), Nf = static_cast<float>(N), Bf = static_cast<float>(B);
  const int32_t* ai = aff_idx + t * K;
  const int32_t* gi = gate_idx != nullptr ? gate_idx + t * K : nullptr;
  const float kb = g3 * static_cast<float>(E) / (Bf * Nf * Nf);
  float qv[C], dq[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int e = lane + 32 * c;
    const bool h = e < E;
    const float pv = h ? p[t * E + e] : 0.f;
    qv[c] = h ? q[t * E + e] : 0.f;
    float cf = g0 / (Tf * E);
    if (in_list(ai, K, e)) cf += g1 / (Tf * K);
    if (gi != nullptr && in_list(gi, K, e)) cf += g2 / (Tf * K);
    if (h) dp[t * E + e] = 2.f * (pv - qv[c]) * cf;
    dq[c] = h ? kb * cnt[b * E + e] : 0.f;
  }
  if (g4 != 0.f) {
    float rv[C], dr[C], dt = 0.f;
    warp_softmax_wide<C>(qv, lane, E, rv);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int e = lane + 32 * c;
      const bool h = e < E;
      const float m = h ? colr[b * E + e] / Nf : 1.f;
      dr[c] = h && m > 0.f ? g4 * (logf(m) + 1.f) / (Bf * Nf) : 0.f;
      dt += rv[c] * dr[c];
    }
    const float dot = warp_sum(dt);
#pragma unroll
    for (int c = 0; c < C; ++c) dq[c] += rv[c] * (dr[c] - dot);
  }
  float dtq = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) dtq += qv[c] * dq[c];
  const float dotq = warp_sum(dtq);
#pragma unroll
  for (int c = 0; c < C; ++c)
    if (lane + 32 * c < E) daff[t * E + lane + 32 * c] = qv[c] * (dq[c] - dotq);
}

inline int n_chunks(int64_t N) { return static_cast<int>((N + kTokPerBlock - 1) / kTokPerBlock); }

}  // namespace
}  // namespace csmoe

using namespace csmoe;

extern "C" int64_t csmoe_losses_workspace_bytes(int64_t B, int64_t N, int32_t E) {
  if (B <= 0 || N <= 0 || E <= 0) return -1;
  if (E > kMaxEWide) return -1;
  const int64_t stride = E > kMaxE ? part_stride(wide_slots(E)) : kPart;
  return B * n_chunks(N) * stride * static_cast<int64_t>(sizeof(float));
}

extern "C" int csmoe_losses_fwd(const float* p, const float* aff, const int32_t* aff_idx, const int32_t* gate_idx, int64_t B,
                                int64_t N, int32_t E, int32_t K, float* q, float* colq, float* cnt, float* colr,
                                float* losses, void* workspace, void* stream_) {
  CSMOE_CHECK_ARG(p && aff && aff_idx && q && colq && cnt && colr && losses && workspace, "csmoe_losses_fwd: NULL argument");
  CSMOE_CHECK_ARG(B >= 1 && N >= 1 && B <= 65535 && E >= 1 && E <= kMaxEWide && K >= 1 && K <= kMaxK && K <= E,
                  "csmoe_losses_fwd: need 1 <= B <= 65535, N >= 1, E <= %d, K <= min(E, %d)", kMaxEWide, kMaxK);
  cudaStream_t stream = as_stream(stream_);
  const int chunks = n_chunks(N);
  float* part = reinterpret_cast<float*>(workspace);
  const dim3 grid(chunks, static_cast<unsigned>(B));
  if (E > 128) {
    losses_stage1_wide<false, 8><<<grid, kWarps * 32, 0, stream>>>(p, aff, aff_idx, gate_idx, N, E, K, q, part);
    CSMOE_CHECK_LAUNCH();
    losses_stage2_wide<false><<<1, 1024, 0, stream>>>(part, 256, chunks, B, N, E, K, colq, cnt, colr, losses);
  } else if (E > kMaxE) {
    losses_stage1_wide<false, 4><<<grid, kWarps * 32, 0, stream>>>(p, aff, aff_idx, gate_idx, N, E, K, q, part);
    CSMOE_CHECK_LAUNCH();
    losses_stage2_wide<false><<<1, 1024, 0, stream>>>(part, 128, chunks, B, N, E, K, colq, cnt, colr, losses);
  } else {
    losses_stage1<false><<<grid, kWarps * 32, 0, stream>>>(p, aff, aff_idx, gate_idx, N, E, K, q, part);
    CSMOE_CHECK_LAUNCH();
    losses_stage2<false><<<1, 1024, 0, stream>>>(part, chunks, B, N, E, K, colq, cnt, colr, losses);
  }
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_losses_bwd(const float* p, const float* q, const int32_t* aff_idx, const int32_t* gate_idx,
                                const float* cnt, const float* colr, const float* g, int64_t B, int64_t N, int32_t E,
                                int32_t K, float* dp, float* daff, void* stream_) {
  CSMOE_CHECK_ARG(p && q && aff_idx && cnt && colr && g && dp && daff, "csmoe_losses_bwd: NULL argument");
  CSMOE_CHECK_ARG(B >= 1 && N >= 1 && E >= 1 && E <= kMaxEWide && K >= 1 && K <= kMaxK, "csmoe_losses_bwd: bad sizes");
  const int64_t T = B * N;
  const unsigned grid = static_cast<unsigned>((T + kWarps - 1) / kWarps);
  if (E > 128)
    losses_bwd_wide_kernel<8><<<grid, kWarps * 32, 0, as_stream(stream_)>>>(p, q, aff_idx, gate_idx, cnt, colr, g, B, N, E, K, dp, daff);
  else if (E > kMaxE)
    losses_bwd_wide_kernel<4><<<grid, kWarps * 32, 0, as_stream(stream_)>>>(p, q, aff_idx, gate_idx, cnt, colr, g, B, N, E, K, dp, daff);
  else
    losses_bwd_kernel<<<grid, kWarps * 32, 0, as_stream(stream_)>>>(p, q, aff_idx, gate_idx, cnt, colr, g, B, N, E, K, dp, daff);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_entropy_balance_fwd(const float* probs, int64_t B, int64_t N, int32_t E, float* colr, float* loss,
                                         void* workspace, void* stream_) {
  CSMOE_CHECK_ARG(probs && colr && loss && workspace, "csmoe_entropy_balance_fwd: NULL argument");
  CSMOE_CHECK_ARG(B >= 1 && N >= 1 && B <= 65535 && E >= 1 && E <= kMaxEWide, "csmoe_entropy_balance_fwd: bad sizes");
  cudaStream_t stream = as_stream(stream_);
  const int chunks = n_chunks(N);
  float* part = reinterpret_cast<float*>(workspace);
  const dim3 grid(chunks, static_cast<unsigned>(B));
  if (E > 128) {
    losses_stage1_wide<true, 8><<<grid, kWarps * 32, 0, stream>>>(probs, nullptr, nullptr, nullptr, N, E, 1, nullptr, part);
    CSMOE_CHECK_LAUNCH();
    losses_stage2_wide<true><<<1, 1024, 0, stream>>>(part, 256, chunks, B, N, E, 1, nullptr, nullptr, colr, loss);
  } else if (E > kMaxE) {
    losses_stage1_wide<true, 4><<<grid, kWarps * 32, 0, stream>>>(probs, nullptr, nullptr, nullptr, N, E, 1, nullptr, part);
    CSMOE_CHECK_LAUNCH();
    losses_stage2_wide<true><<<1, 1024, 0, stream>>>(part, 128, chunks, B, N, E, 1, nullptr, nullptr, colr, loss);
  } else {
    losses_stage1<true><<<grid, kWarps * 32, 0, stream>>>(probs, nullptr, nullptr, nullptr, N, E, 1, nullptr, part);
    CSMOE_CHECK_LAUNCH();
    losses_stage2<true><<<1, 1024, 0, stream>>>(part, chunks, B, N, E, 1, nullptr, nullptr, colr, loss);
  }
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_entropy_balance_bwd(const float* colr, const float* g, int64_t B, int64_t N, int32_t E, float* dprobs,
                                         void* stream_) {
  CSMOE_CHECK_ARG(colr && g && dprobs && B >= 1 && N >= 1 && E >= 1, "csmoe_entropy_balance_bwd: bad arguments");
  const int64_t n = B * N * E;
  entropy_balance_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream_)>>>(colr, g, B, N, E, dprobs);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_topk_renorm_bwd(const float* scores, const float* w, const int32_t* idx, const float* dw, int64_t T,
                                     int32_t E, int32_t K, int32_t sigmoid, int32_t accumulate, float* dscores,
                                     void* stream_) {
  CSMOE_CHECK_ARG(scores && w && idx && dw && dscores, "csmoe_topk_renorm_bwd: NULL argument");
  CSMOE_CHECK_ARG(T >= 0 && E >= 1 && K >= 1 && K <= kMaxK && K <= E, "csmoe_topk_renorm_bwd: bad sizes");
  if (T == 0) return CSMOE_OK;
  topk_renorm_bwd_kernel<<<static_cast<unsigned>((T + 127) / 128), 128, 0, as_stream(stream_)>>>(
      scores, w, idx, dw, T, E, K, sigmoid, accumulate, dscores);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_dense_rows(const int32_t* idx, int64_t T, int32_t K, int64_t t_pad, int32_t* rows, void* stream_) {
  CSMOE_CHECK_ARG(idx && rows && T >= 0 && K >= 1 && t_pad >= T, "csmoe_dense_rows: bad arguments");
  if (T == 0) return CSMOE_OK;
  dense_rows_kernel<<<static_cast<unsigned>((T * K + 255) / 256), 256, 0, as_stream(stream_)>>>(idx, T, K, t_pad, rows);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}
