// Router: skinny gate GEMM fused with fp32 softmax, warp-shuffle top-k and renormalisation (one warp per token).
// Replaces router_policy + topk_expert (moe_model/model/moe/competesmoe.py:301-320, moe.py:113-132;
// moe_pretrain_model/layers/moe/competesmoe.py:465-490, moe.py:373-393).
//
// Rounding points follow the reference: logits are accumulated in fp32 and rounded to the activation dtype (the
// nn.Linear / F.linear output), softmax runs in fp32 on the rounded logits, the top-k weights stay fp32 and are divided
// by their sum rounded to the dtype of the LAYER INPUT (`.to(x.dtype)` on the denominator only): `renorm_dtype`, which is
// the activation dtype for a bf16 model (multimodal plugin) and fp32 for fp32 inputs under autocast (pretrain plugin).
// Tie-break: highest value first, equal values -> lowest expert index (the reference's torch.topk is unspecified on
// ties; see DESIGN.md "routing parity").
#include "common.h"

namespace csmoe {
namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxE = 64;  // two candidates per lane
constexpr int kMaxK = 8;

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

// Total order on floats for the selection: an order-preserving integer key, NaN (either sign) sorted as the largest
// value, which is where torch.topk puts it.  Comparisons on the key are total, so all lanes always agree on the winner
// and a token with non-finite scores still gets K distinct, in-range expert ids (the loss then goes NaN exactly like the
// reference's; nothing indexes out of bounds downstream).
__device__ __forceinline__ unsigned order_key(float v) {
  const unsigned b = __float_as_uint(v);
  if ((b & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;   // NaN
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Warp-wide selection of the K largest of up to 64 values (lane l holds experts l and l+32).
// Emits (value, index) pairs in descending order to lane-uniform arrays.  Ties: lowest expert index first.
// Requires K <= E (checked on the host): then every round has an untaken valid candidate and the sentinel never wins.
__device__ __forceinline__ void warp_topk(float v0, float v1, int lane, int E, int K, float (&out_v)[kMaxK],
                                          int (&out_i)[kMaxK]) {
  bool t0 = lane >= E, t1 = lane + 32 >= E;  // taken / invalid
  const unsigned k0 = order_key(v0), k1 = order_key(v1);
#pragma unroll
  for (int k = 0; k < kMaxK; ++k) {
    if (k >= K) break;
    unsigned bk;
    int bi;
    // local best (lower index wins ties): candidate 0 has the lower index
    if (!t0 && (t1 || k0 >= k1)) {
      bk = k0;
      bi = lane;
    } else if (!t1) {
      bk = k1;
      bi = lane + 32;
    } else {
      bk = 0u;            // below the key of every valid value (-inf maps to 0x007fffff)
      bi = 0x7fffffff;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned ok = __shfl_xor_sync(0xffffffffu, bk, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ok > bk || (ok == bk && oi < bi)) {
        bk = ok;
        bi = oi;
      }
    }
    if (bi == 0x7fffffff) bi = k < E ? k : 0;   // unreachable for K <= E; keeps the index in range regardless
    out_v[k] = __shfl_sync(0xffffffffu, bi >= 32 ? v1 : v0, bi & 31);
    out_i[k] = bi;
    if (bi == lane) t0 = true;
    if (bi == lane + 32) t1 = true;
  }
}

template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
router_fwd_kernel(const T* __restrict__ x, const T* __restrict__ wg, long long Tn, int D, int E, int K, int renorm_bf16,
                  T* __restrict__ logits, float* __restrict__ probs, float* __restrict__ topk_w,
                  int32_t* __restrict__ topk_idx) {
  const int lane = threadIdx.x & 31;
  const long long t = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (t >= Tn) return;
  const T* xr = x + t * D;
  float l0 = 0.f, l1 = 0.f;  // logits of experts `lane` and `lane + 32`
  constexpr int EC = 4;
  for (int e0 = 0; e0 < E; e0 += EC) {
    float acc[EC] = {0.f, 0.f, 0.f, 0.f};
    for (int d0 = lane * 8; d0 < D; d0 += 1024) {     // four 16-byte x loads in flight per lane (same summation order)
      float xv[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (d0 + u * 256 < D) load8(xr + d0 + u * 256, xv[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int d = d0 + u * 256;
        if (d < D) {
#pragma unroll
          for (int i = 0; i < EC; ++i) {
            if (e0 + i < E) {
              float wv[8];
              load8(wg + static_cast<long long>(e0 + i) * D + d, wv);
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[i] = fmaf(xv[u][j], wv[j], acc[i]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < EC; ++i) {
      const float s = round_as(warp_sum(acc[i]), static_cast<const T*>(nullptr));
      const int e = e0 + i;
      if (e < E) {
        if (e == lane) l0 = s;
        if (e == lane + 32) l1 = s;
      }
    }
  }
  const bool has0 = lane < E, has1 = lane + 32 < E;
  float m = warp_max(fmaxf(has0 ? l0 : -INFINITY, has1 ? l1 : -INFINITY));
  const float e0v = has0 ? expf(l0 - m) : 0.f, e1v = has1 ? expf(l1 - m) : 0.f;
  const float denom = warp_sum(e0v + e1v);
  const float p0 = e0v / denom, p1 = e1v / denom;
  if (has0) {
    logits[t * E + lane] = static_cast<T>(l0);
    probs[t * E + lane] = p0;
  }
  if (has1) {
    logits[t * E + lane + 32] = static_cast<T>(l1);
    probs[t * E + lane + 32] = p1;
  }
  float tv[kMaxK];
  int ti[kMaxK];
  warp_topk(p0, p1, lane, E, K, tv, ti);
  if (lane == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxK; ++k)
      if (k < K) s += tv[k];
    if (renorm_bf16) s = bf16_round(s);
#pragma unroll
    for (int k = 0; k < kMaxK; ++k)
      if (k < K) {
        topk_w[t * K + k] = tv[k] / s;
        topk_idx[t * K + k] = ti[k];
      }
  }
}

// Softmax + top-k + renormalisation from logits that a tensor-core GEMM already produced (many experts: the skinny
// CUDA-core dot products of router_fwd_kernel cost 64 warp reductions per token at E = 64).  Same outputs and rounding
// points as router_fwd_kernel; logits are read in the activation dtype.
template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
router_from_logits_kernel(const T* __restrict__ logits, long long Tn, int E, int K, int renorm_bf16, float* __restrict__ probs,
                          float* __restrict__ topk_w, int32_t* __restrict__ topk_idx) {
  const int lane = threadIdx.x & 31;
  const long long t = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (t >= Tn) return;
  const bool has0 = lane < E, has1 = lane + 32 < E;
  const float l0 = has0 ? static_cast<float>(logits[t * E + lane]) : -INFINITY;
  const float l1 = has1 ? static_cast<float>(logits[t * E + lane + 32]) : -INFINITY;
  const float m = warp_max(fmaxf(l0, l1));
  const float e0v = has0 ? expf(l0 - m) : 0.f, e1v = has1 ? expf(l1 - m) : 0.f;
  const float denom = warp_sum(e0v + e1v);
  const float p0 = e0v / denom, p1 = e1v / denom;
  if (has0) probs[t * E + lane] = p0;
  if (has1) probs[t * E + lane + 32] = p1;
  float tv[kMaxK];
  int ti[kMaxK];
  warp_topk(p0, p1, lane, E, K, tv, ti);
  if (lane == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxK; ++k)
      if (k < K) s += tv[k];
    if (renorm_bf16) s = bf16_round(s);
#pragma unroll
    for (int k = 0; k < kMaxK; ++k)
      if (k < K) {
        topk_w[t * K + k] = tv[k] / s;
        topk_idx[t * K + k] = ti[k];
      }
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
topk_renorm_kernel(const float* __restrict__ scores, long long Tn, int E, int K, int mode, int round_dtype,
                   float* __restrict__ topk_w, int32_t* __restrict__ topk_idx) {
  const int lane = threadIdx.x & 31;
  const long long t = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (t >= Tn) return;
  const bool sigmoid = (mode & 1) != 0, round_out = (mode & 2) != 0, bf = round_dtype == CSMOE_BF16;
  float v0 = lane < E ? scores[t * E + lane] : 0.f;
  float v1 = lane + 32 < E ? scores[t * E + lane + 32] : 0.f;
  if (sigmoid) {
    v0 = 1.f / (1.f + expf(-v0));
    v1 = 1.f / (1.f + expf(-v1));
    if (bf) {
      v0 = bf16_round(v0);
      v1 = bf16_round(v1);
    }
  }
  float tv[kMaxK];
  int ti[kMaxK];
  warp_topk(v0, v1, lane, E, K, tv, ti);
  if (lane == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxK; ++k)
      if (k < K) s += tv[k];
    if (bf) s = bf16_round(s);
#pragma unroll
    for (int k = 0; k < kMaxK; ++k)
      if (k < K) {
        float w = tv[k] / s;
        if (round_out && bf) w = bf16_round(w);
        topk_w[t * K + k] = w;
        topk_idx[t * K + k] = ti[k];
      }
  }
}


// ------------------------------------------------------------------------------------------------ aux losses
// Balance loss (moe.py:90-110) and z-loss (moe.py:71-88) of the router step need, per batch element b and expert e,
// S[b,e] = sum_n p[b,n,e] and C[b,e] = #{n : top-1(b,n) = e}, plus Z = sum_t logsumexp(logits_t)^2.
// Stage 1: one CTA per (b, token chunk) writes partial sums; stage 2 (one CTA) adds them in a fixed order
// (deterministic) and evaluates   balance = E^2 * mean_{b,e}(S/N * C/N),   z = Z / T.
constexpr int kAuxChunk = 256;  // tokens per stage-1 CTA (8 warps x 32 tokens)

template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
router_aux_stage1(const T* __restrict__ logits, const float* __restrict__ probs, const int32_t* __restrict__ topk_idx,
                  int N, int E, int K, int chunks_per_b, float* __restrict__ partial, float* __restrict__ lse_out) {
  extern __shared__ float sh[];  // [8 warps][2E + 1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / chunks_per_b, c = blockIdx.x % chunks_per_b;
  const int stride = 2 * E + 1;
  float s0 = 0.f, s1 = 0.f, c0 = 0.f, c1 = 0.f, zz = 0.f;  // lane l: experts l and l+32
  const int n0 = c * kAuxChunk + warp * 32;
  if (E <= 8) {
    // few experts: one token per lane (the expert-per-lane loop below would walk 32 tokens serially with E lanes busy)
    const int n = n0 + lane;
    const bool ok = n < N;
    const long long t = static_cast<long long>(b) * N + (ok ? n : 0);
    float lg[8], m = -INFINITY;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      lg[e] = (ok && e < E) ? static_cast<float>(logits[t * E + e]) : -INFINITY;
      m = fmaxf(m, lg[e]);
    }
    float se = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (e < E) se += expf(lg[e] - m);
    const float lse = ok ? m + logf(se) : 0.f;
    if (ok && lse_out) lse_out[t] = lse;
    zz = warp_sum(lse * lse);
    const int top1 = ok ? topk_idx[t * K] : -1;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (e < E) {
        const float ps = warp_sum(ok ? probs[t * E + e] : 0.f);
        const float cs = warp_sum(top1 == e ? 1.f : 0.f);
        if (lane == e) { s0 = ps; c0 = cs; }
      }
    }
  } else
  for (int i = 0; i < 32; ++i) {
    const int n = n0 + i;
    if (n >= N) break;
    const long long t = static_cast<long long>(b) * N + n;
    const bool h0 = lane < E, h1 = lane + 32 < E;
    const float l0 = h0 ? static_cast<float>(logits[t * E + lane]) : -INFINITY;
    const float l1 = h1 ? static_cast<float>(logits[t * E + lane + 32]) : -INFINITY;
    const float m = warp_max(fmaxf(l0, l1));
    const float se = warp_sum((h0 ? expf(l0 - m) : 0.f) + (h1 ? expf(l1 - m) : 0.f));
    const float lse = m + logf(se);
    if (lane == 0 && lse_out) lse_out[t] = lse;
    zz += lse * lse;
    if (h0) s0 += probs[t * E + lane];
    if (h1) s1 += probs[t * E + lane + 32];
    const int top1 = topk_idx[t * K];
    if (top1 == lane) c0 += 1.f;
    if (top1 == lane + 32) c1 += 1.f;
  }
  float* mine = sh + warp * stride;
  if (lane < E) { mine[lane] = s0; mine[E + lane] = c0; }
  if (lane + 32 < E) { mine[lane + 32] = s1; mine[E + lane + 32] = c1; }
  if (lane == 0) mine[2 * E] = zz;
  __syncthreads();
  for (int i = threadIdx.x; i < stride; i += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) a += sh[w * stride + i];
    partial[static_cast<long long>(blockIdx.x) * stride + i] = a;
  }
}

__global__ void __launch_bounds__(256)
router_aux_stage2(const float* __restrict__ partial, int B, int N, int E, int chunks_per_b, float* __restrict__ psum,
                  float* __restrict__ cnt, float* __restrict__ losses) {
  __shared__ float red[256];
  const int stride = 2 * E + 1;
  float bal = 0.f, zz = 0.f;
  for (int i = threadIdx.x; i < B * E; i += blockDim.x) {
    const int b = i / E, e = i % E;
    float s = 0.f, c = 0.f;
    for (int k = 0; k < chunks_per_b; ++k) {
      const float* p = partial + static_cast<long long>(b * chunks_per_b + k) * stride;
      s += p[e];
      c += p[E + e];
    }
    psum[i] = s;
    cnt[i] = c;
    bal += (s / N) * (c / N);
  }
  for (int i = threadIdx.x; i < B * chunks_per_b; i += blockDim.x) zz += partial[static_cast<long long>(i) * stride + 2 * E];
  red[threadIdx.x] = bal;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const float bal_tot = red[0];
  __syncthreads();
  red[threadIdx.x] = zz;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    losses[0] = bal_tot / (static_cast<float>(B) * E) * static_cast<float>(E) * E;
    losses[1] = red[0] / (static_cast<float>(B) * N);
  }
}

// ------------------------------------------------------------------------------------------------ router backward
// Stage A (one warp per token): total gradient w.r.t. the gate logits,
//   dp  = dprobs_in + g_bal * E/(B N^2) * C[b,e] + scatter_k( dtw_k / r + round(-sum_j dtw_j (w_j / r)) )
//         r = round(sum_k p_{i_k}), round = to `renorm_dtype`: autograd's gradients of `w / sum(w).to(x.dtype)` -- the
//         quotient's numerator gradient in fp32, the denominator's accumulated and then cast to the denominator's dtype
//         (with an fp32 denominator this is (dtw_k - sum_j dtw_j w_j) / s)
//   dl  = p * (dp - <dp, p>) + dlogits_in + g_z * (2 lse / T) * p
// rounded to the activation dtype (the reference back-propagates through a bf16 nn.Linear), then dx[t,:] = dl . Wg.
template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
router_bwd_dx_kernel(const T* __restrict__ wg, const float* __restrict__ probs, const float* __restrict__ topk_w,
                     const int32_t* __restrict__ topk_idx, const float* __restrict__ dtw, const float* __restrict__ dprobs_in,
                     const float* __restrict__ dlogits_in, const float* __restrict__ lse, const float* __restrict__ cnt,
                     const float* __restrict__ g_losses, long long Tn, int N, int B, int D, int E, int K, int renorm_bf16,
                     float* __restrict__ dl_out, T* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const long long t = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (t >= Tn) return;
  const bool h0 = lane < E, h1 = lane + 32 < E;
  const float p0 = h0 ? probs[t * E + lane] : 0.f, p1 = h1 ? probs[t * E + lane + 32] : 0.f;
  float d0 = (h0 && dprobs_in) ? dprobs_in[t * E + lane] : 0.f;
  float d1 = (h1 && dprobs_in) ? dprobs_in[t * E + lane + 32] : 0.f;
  if (g_losses != nullptr && cnt != nullptr) {
    const int b = static_cast<int>(t / N);
    const float coef = g_losses[0] * static_cast<float>(E) / (static_cast<float>(B) * N * N);
    if (h0) d0 += coef * cnt[b * E + lane];
    if (h1) d1 += coef * cnt[b * E + lane + 32];
  }
  if (dtw != nullptr) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += probs[t * E + topk_idx[t * K + k]];
    if (renorm_bf16) s = bf16_round(s);
    float dden = 0.f;
    for (int k = 0; k < K; ++k) dden -= dtw[t * K + k] * (topk_w[t * K + k] / s);
    if (renorm_bf16) dden = bf16_round(dden);
    for (int k = 0; k < K; ++k) {
      const int i = topk_idx[t * K + k];
      const float g = dtw[t * K + k] / s + dden;
      if (i == lane) d0 += g;
      if (i == lane + 32) d1 += g;
    }
  }
  const float inner = warp_sum(d0 * p0 + d1 * p1);
  float l0 = p0 * (d0 - inner), l1 = p1 * (d1 - inner);
  if (dlogits_in != nullptr) {
    if (h0) l0 += dlogits_in[t * E + lane];
    if (h1) l1 += dlogits_in[t * E + lane + 32];
  }
  if (g_losses != nullptr && lse != nullptr) {
    const float gz = g_losses[1] * 2.f * lse[t] / static_cast<float>(Tn);
    l0 += gz * p0;
    l1 += gz * p1;
  }
  l0 = round_as(l0, static_cast<const T*>(nullptr));
  l1 = round_as(l1, static_cast<const T*>(nullptr));
  if (h0) dl_out[t * E + lane] = l0;
  if (h1) dl_out[t * E + lane + 32] = l1;
  if (dx == nullptr) return;
  for (int d0 = 0; d0 < D; d0 += 256) {     // every lane runs every pass: the shuffle reads dl from lane e & 31, which
    const int d = d0 + lane * 8;            // must be there also when that lane has no columns left (D % 256 != 0)
    const bool on = d < D;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int e = 0; e < E; ++e) {
      const float g = __shfl_sync(0xffffffffu, e < 32 ? l0 : l1, e & 31);
      if (on) {
        float wv[8];
        load8(wg + static_cast<long long>(e) * D + d, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(g, wv[j], acc[j]);
      }
    }
    if (on) store8(dx + t * D + d, acc);
  }
}

// Stage B: partial dWg[e, :] = sum_{t in chunk} dl[t,e] * x[t,:]  (8 experts per pass), stage C adds the chunks in order.
constexpr int kDwChunk = 32;   // tokens per stage-1 CTA: T/32 x D/1024 CTAs keep the SMs busy even for E = 4
constexpr int kDwExperts = 8;

template <typename T>
__global__ void __launch_bounds__(128)
router_bwd_dw_stage1(const T* __restrict__ x, const float* __restrict__ dl, long long Tn, int D, int E,
                     float* __restrict__ partial) {
  __shared__ float sdl[kDwChunk][kDwExperts];
  const int e0 = blockIdx.z * kDwExperts;
  const long long t0 = static_cast<long long>(blockIdx.y) * kDwChunk;
  const int col = (blockIdx.x * 128 + threadIdx.x) * 8;
  const int nt = static_cast<int>(min(static_cast<long long>(kDwChunk), Tn - t0));
  for (int i = threadIdx.x; i < kDwChunk * kDwExperts; i += 128) {
    const int tt = i / kDwExperts, ee = i % kDwExperts;
    sdl[tt][ee] = (tt < nt && e0 + ee < E) ? dl[(t0 + tt) * E + e0 + ee] : 0.f;
  }
  __syncthreads();
  if (col >= D) return;
  float acc[kDwExperts][8];
#pragma unroll
  for (int e = 0; e < kDwExperts; ++e)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[e][j] = 0.f;
  for (int t4 = 0; t4 < nt; t4 += 4) {            // four token rows in flight per thread; summation order unchanged
    float xv[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (t4 + u < nt) load8(x + (t0 + t4 + u) * D + col, xv[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (t4 + u < nt) {
#pragma unroll
        for (int e = 0; e < kDwExperts; ++e) {
          const float g = sdl[t4 + u][e];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[e][j] = fmaf(g, xv[u][j], acc[e][j]);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < kDwExperts; ++e) {
    if (e0 + e < E) store8(partial + (static_cast<long long>(blockIdx.y) * E + e0 + e) * D + col, acc[e]);
  }
}

// dWg[i] = sum over chunks of partial[c][i] in a fixed order: 32 chunk lanes x 8 column vectors per block, lane l adds
// chunks l, l+32, ... and lane 0 adds the 32 lane sums in lane order (deterministic; the earlier one-thread-per-vector
// loop walked all T/32 chunks serially: 55 us at T = 12800).
template <typename WT>
__global__ void __launch_bounds__(256)
router_bwd_dw_stage2(const float* __restrict__ partial, int n_chunks, long long ED, WT* __restrict__ dwg) {
  __shared__ float red[32][8][8];
  const int v = threadIdx.x & 7, cl = threadIdx.x >> 3;
  const long long i = (static_cast<long long>(blockIdx.x) * 8 + v) * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (i < ED) {
    for (int c = cl; c < n_chunks; c += 32) {
      float t[8];
      load8(partial + c * ED + i, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += t[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[cl][v][j] = acc[j];
  __syncthreads();
  if (cl == 0 && i < ED) {
#pragma unroll 1
    for (int l = 1; l < 32; ++l)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += red[l][v][j];
    store8(dwg + i, acc);
  }
}

// ------------------------------------------------------------------------------------------------ more than 64 experts
// The kernels above keep two experts per lane (E <= 64: every shipped sweep, and the shapes that were measured).  The
// pretrain plugin's own default is -moe.n_experts 128 (transformer_lm_mixin.py:32), so for 64 < E <= 256 the same
// algorithms run with C = 4 or 8 experts per lane: lane l holds experts l, l + 32, ..., l + 32 (C - 1).  Rounding points,
// the order of every reduction and the tie-break (highest value, then lowest expert index) are the ones stated above.
constexpr int kMaxEWide = 256;

template <int C>
__device__ __forceinline__ float slot_of(const float (&v)[C], int c) {   // v[c] for a warp-uniform c, without local memory
  float r = v[0];
#pragma unroll
  for (int i = 1; i < C; ++i)
    if (c == i) r = v[i];
  return r;
}

template <int C>
__device__ __forceinline__ void warp_topk_wide(const float (&v)[C], int lane, int E, int K, float (&out_v)[kMaxK],
                                               int (&out_i)[kMaxK]) {
  bool taken[C];
  unsigned key[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    taken[c] = lane + 32 * c >= E;
    key[c] = order_key(v[c]);
  }
#pragma unroll
  for (int k = 0; k < kMaxK; ++k) {
    if (k >= K) break;
    unsigned bk = 0u;            // below the key of every valid value
    int bi = 0x7fffffff;
#pragma unroll
    for (int c = 0; c < C; ++c)  // slots in ascending expert order: a later one wins only when strictly larger
      if (!taken[c] && (bi == 0x7fffffff || key[c] > bk)) {
        bk = key[c];
        bi = lane + 32 * c;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned ok = __shfl_xor_sync(0xffffffffu, bk, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ok > bk || (ok == bk && oi < bi)) {
        bk = ok;
        bi = oi;
      }
    }
    if (bi == 0x7fffffff) bi = k < E ? k : 0;   // unreachable for K <= E
    out_v[k] = __shfl_sync(0xffffffffu, slot_of<C>(v, bi >> 5), bi & 31);
    out_i[k] = bi;
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (bi == lane + 32 * c) taken[c] = true;
  }
}

// softmax of the C values per lane (invalid slots excluded); returns the probabilities in place
template <int C>
__device__ __forceinline__ void warp_softmax_wide(float (&l)[C], int lane, int E) {
  float m = -INFINITY;
#pragma unroll
  for (int c = 0; c < C; ++c)
    if (lane + 32 * c < E) m = fmaxf(m, l[c]);
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    l[c] = lane + 32 * c < E ? expf(l[c] - m) : 0.f;
    s += l[c];
  }
  const float denom = warp_sum(s);
#pragma unroll
  for (int c = 0; c < C; ++c) l[c] = l[c] / denom;
}

__device__ __forceinline__ void write_topk(const float (&tv)[kMaxK], const int (&ti)[kMaxK], int K, int renorm_bf16, long long t,
                                           float* __restrict__ topk_w, int32_t* __restrict__ topk_idx) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxK; ++k)
    if (k < K) s += tv[k];
  if (renorm_bf16) s = bf16_round(s);
#pragma unroll
  for (int k = 0; k < kMaxK; ++k)
    if (k < K) {
      topk_w[t * K + k] = tv[k] / s;
      topk_idx[t * K + k] = ti[k];
    }
}

template <typename T, int C>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
router_fwd_wide_kernel(const T* __restrict__ x, const T* __restrict__ wg, long long Tn, int D, int E, int K, int renorm_bf16,
                       T* __restrict__ logits, float* __restrict__ probs, float* __restrict__ topk_w,
                       int32_t* __restrict__ topk_idx) {
  const int lane = threadIdx.x & 31;
  const long long t = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (t >= Tn) return;
  const T* xr = x + t * D;
  float l[C];
#pragma unroll
  for (int c = 0; c < C; ++c) l[c] = 0.f;
  constexpr int EC = 4;
  for (int e0 = 0; e0 < E; e0 += EC) {
    float acc[EC] = {0.f, 0.f, 0.f, 0.f};
    for (int d = lane * 8; d < D; d += 256) {
      float xv[8];
      load8(xr + d, xv);
#pragma unroll
      for (int i = 0; i < EC; ++i) {
        if (e0 + i < E) {
          float wv[8];
          load8(wg + static_cast<long long>(e0 + i) * D + d, wv);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i] = fmaf(xv[j], wv[j], acc[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < EC; ++i) {
      const float s = round_as(warp_sum(acc[i]), static_cast<const T*>(nullptr));
      const int e = e0 + i;
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (e < E && e == lane + 32 * c) l[c] = s;
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c)
    if (lane + 32 * c < E) logits[t * E + lane + 32 * c] = static_cast<T>(l[c]);
  warp_softmax_wide<C>(l, lane, E);
#pragma unroll
  for (int c = 0; c < C; ++c)
    if (lane + 32 * c < E) probs[t * E + lane + 32 * c] = l[c];
  float tv[kMaxK];
  int ti[kMaxK];
  warp_topk_wide<C>(l, lane, E, K, tv, ti);
  if (lane == 0) write_topk(tv, ti, K, renorm_bf16, t, topk_w, topk_idx);
}

template <typename T, int C>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
router_from_logits_wide_kernel(const T* __restrict__ logits, long long Tn, int E, int K, int renorm_bf16, float* __restrict__ probs,
                               float* __restrict__ topk_w, int32_t* __restrict__ topk_idx) {
  const int lane = threadIdx.x & 31;
  const long long t = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (t >= Tn) return;
  float l[C];
#pragma unroll
  for (int c = 0; c < C; ++c) l[c] = lane + 32 * c < E ? static_cast<float>(logits[t * E + lane + 32 * c]) : -INFINITY;
  warp_softmax_wide<C>(l, lane, E);
#pragma unroll
  for (int c = 0; c < C; ++c)
    if (lane + 32 * c < E) probs[t * E + lane + 32 * c] = l[c];
  float tv[kMaxK];
  int ti[kMaxK];
  warp_topk_wide<C>(l, lane, E, K, tv, ti);
  if (lane == 0) write_topk(tv, ti, K, renorm_bf16, t, topk_w, topk_idx);
}

template <int C>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
topk_renorm_wide_kernel(const float* __restrict__ scores, long long Tn, int E, int K, int mode, int round_dtype,
                        float* __restrict__ topk_w, int32_t* __restrict__ topk_idx) {
  const int lane = threadIdx.x & 31;
  const long long t = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (t >= Tn) return;
  const bool sigmoid = (mode & 1) != 0, round_out = (mode & 2) != 0, bf = round_dtype == CSMOE_BF16;
  float v[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    v[c] = lane + 32 * c < E ? scores[t * E + lane + 32 * c] : 0.f;
    if (sigmoid) {
      v[c] = 1.f / (1.f + expf(-v[c]));
      if (bf) v[c] = bf16_round(v[c]);
    }
  }
  float tv[kMaxK];
  int ti[kMaxK];
  warp_topk_wide<C>(v, lane, E, K, tv, ti);
  if (lane == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxK; ++k)
      if (k < K) s += tv[k];
    if (bf) s = bf16_round(s);
#pragma unroll
    for (int k = 0; k < kMaxK; ++k)
      if (k < K) {
        float w = tv[k] / s;
        if (round_out && bf) w = bf16_round(w);
        topk_w[t * K + k] = w;
        topk_idx[t * K + k] = ti[k];
      }
  }
}

// stage 1 of the balance / z losses (same partial layout as router_aux_stage1: [chunk][2E + 1]); stage 2 is shared
template <typename T, int C>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
router_aux_stage1_wide(const T* __restrict__ logits, const float* __restrict__ probs, const int32_t* __restrict__ topk_idx,
                       int N, int E, int K, int chunks_per_b, float* __restrict__ partial, float* __restrict__ lse_out) {
  extern __shared__ float sh[];  // [8 warps][2E + 1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / chunks_per_b, ch = blockIdx.x % chunks_per_b;
  const int stride = 2 * E + 1;
  float sp[C], sc[C], zz = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) sp[c] = sc[c] = 0.f;
  const int n0 = ch * kAuxChunk + warp * 32;
  for (int i = 0; i < 32; ++i) {
    const int n = n0 + i;
    if (n >= N) break;
    const long long t = static_cast<long long>(b) * N + n;
    float m = -INFINITY, lg[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      lg[c] = lane + 32 * c < E ? static_cast<float>(logits[t * E + lane + 32 * c]) : -INFINITY;
      m = fmaxf(m, lg[c]);
    }
    m = warp_max(m);
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) se += lane + 32 * c < E ? expf(lg[c] - m) : 0.f;
    se = warp_sum(se);
    const float lse = m + logf(se);
    if (lane == 0 && lse_out) lse_out[t] = lse;
    zz += lse * lse;
    const int top1 = topk_idx[t * K];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (lane + 32 * c < E) sp[c] += probs[t * E + lane + 32 * c];
      if (top1 == lane + 32 * c) sc[c] += 1.f;
    }
  }
  float* mine = sh + warp * stride;
#pragma unroll
  for (int c = 0; c < C; ++c)
    if (lane + 32 * c < E) {
      mine[lane + 32 * c] = sp[c];
      mine[E + lane + 32 * c] = sc[c];
    }
  if (lane == 0) mine[2 * E] = zz;
  __syncthreads();
  for (int i = threadIdx.x; i < stride; i += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) a += sh[w * stride + i];
    partial[static_cast<long long>(blockIdx.x) * stride + i] = a;
  }
}

template <typename T, int C>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
router_bwd_dx_wide_kernel(const T* __restrict__ wg, const float* __restrict__ probs, const float* __restrict__ topk_w,
                          const int32_t* __restrict__ topk_idx, const float* __restrict__ dtw,
                          const float* __restrict__ dprobs_in, const float* __restrict__ dlogits_in,
                          const float* __restrict__ lse, const float* __restrict__ cnt, const float* __restrict__ g_losses,
                          long long Tn, int N, int B, int D, int E, int K, int renorm_bf16, float* __restrict__ dl_out,
                          T* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const long long t = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (t >= Tn) return;
  float p[C], d[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const bool h = lane + 32 * c < E;
    p[c] = h ? probs[t * E + lane + 32 * c] : 0.f;
    d[c] = (h && dprobs_in) ? dprobs_in[t * E + lane + 32 * c] : 0.f;
  }
  if (g_losses != nullptr && cnt != nullptr) {
    const int b = static_cast<int>(t / N);
    const float coef = g_losses[0] * static_cast<float>(E) / (static_cast<float>(B) * N * N);
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (lane + 32 * c < E) d[c] += coef * cnt[b * E + lane + 32 * c];
  }
  if (dtw != nullptr) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += probs[t * E + topk_idx[t * K + k]];
    if (renorm_bf16) s = bf16_round(s);
    float dden = 0.f;
    for (int k = 0; k < K; ++k) dden -= dtw[t * K + k] * (topk_w[t * K + k] / s);
    if (renorm_bf16) dden = bf16_round(dden);
    for (int k = 0; k < K; ++k) {
      const int i = topk_idx[t * K + k];
      const float g = dtw[t * K + k] / s + dden;
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (i == lane + 32 * c) d[c] += g;
    }
  }
  float in = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) in += d[c] * p[c];
  const float inner = warp_sum(in);
  const float gz = (g_losses != nullptr && lse != nullptr) ? g_losses[1] * 2.f * lse[t] / static_cast<float>(Tn) : 0.f;
  float l[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const bool h = lane + 32 * c < E;
    l[c] = p[c] * (d[c] - inner);
    if (dlogits_in != nullptr && h) l[c] += dlogits_in[t * E + lane + 32 * c];
    if (g_losses != nullptr && lse != nullptr) l[c] += gz * p[c];
    l[c] = round_as(l[c], static_cast<const T*>(nullptr));
    if (h) dl_out[t * E + lane + 32 * c] = l[c];
  }
  if (dx == nullptr) return;
  for (int d0 = 0; d0 < D; d0 += 256) {     // every lane runs every pass (the shuffle needs the whole warp)
    const int dd = d0 + lane * 8;
    const bool on = dd < D;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int e = 0; e < E; ++e) {
      const float g = __shfl_sync(0xffffffffu, slot_of<C>(l, e >> 5), e & 31);
      if (on) {
        float wv[8];
        load8(wg + static_cast<long long>(e) * D + dd, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(g, wv[j], acc[j]);
      }
    }
    if (on) store8(dx + t * D + dd, acc);
  }
}

// ---- host side: two experts per lane up to 64 experts, C = 4 up to 128, C = 8 up to 256
template <typename T>
void launch_router_fwd(unsigned grid, cudaStream_t stream, const T* x, const T* wg, long long Tn, int D, int E, int K, int rbf,
                       T* logits, float* probs, float* topk_w, int32_t* topk_idx) {
  if (E <= kMaxE)
    router_fwd_kernel<T><<<grid, kWarpsPerBlock * 32, 0, stream>>>(x, wg, Tn, D, E, K, rbf, logits, probs, topk_w, topk_idx);
  else if (E <= 128)
    router_fwd_wide_kernel<T, 4><<<grid, kWarpsPerBlock * 32, 0, stream>>>(x, wg, Tn, D, E, K, rbf, logits, probs, topk_w, topk_idx);
  else
    router_fwd_wide_kernel<T, 8><<<grid, kWarpsPerBlock * 32, 0, stream>>>(x, wg, Tn, D, E, K, rbf, logits, probs, topk_w, topk_idx);
}

template <typename T>
void launch_router_from_logits(unsigned grid, cudaStream_t stream, const T* logits, long long Tn, int E, int K, int rbf,
                               float* probs, float* topk_w, int32_t* topk_idx) {
  if (E <= kMaxE)
    router_from_logits_kernel<T><<<grid, kWarpsPerBlock * 32, 0, stream>>>(logits, Tn, E, K, rbf, probs, topk_w, topk_idx);
  else if (E <= 128)
    router_from_logits_wide_kernel<T, 4><<<grid, kWarpsPerBlock * 32, 0, stream>>>(logits, Tn, E, K, rbf, probs, topk_w, topk_idx);
  else
    router_from_logits_wide_kernel<T, 8><<<grid, kWarpsPerBlock * 32, 0, stream>>>(logits, Tn, E, K, rbf, probs, topk_w, topk_idx);
}

template <typename T>
void launch_router_aux_stage1(unsigned grid, size_t smem, cudaStream_t stream, const T* logits, const float* probs,
                              const int32_t* topk_idx, int N, int E, int K, int chunks, float* partial, float* lse) {
  if (E <= kMaxE)
    router_aux_stage1<T><<<grid, kWarpsPerBlock * 32, smem, stream>>>(logits, probs, topk_idx, N, E, K, chunks, partial, lse);
  else if (E <= 128)
    router_aux_stage1_wide<T, 4><<<grid, kWarpsPerBlock * 32, smem, stream>>>(logits, probs, topk_idx, N, E, K, chunks, partial, lse);
  else
    router_aux_stage1_wide<T, 8><<<grid, kWarpsPerBlock * 32, smem, stream>>>(logits, probs, topk_idx, N, E, K, chunks, partial, lse);
}

template <typename T>
void launch_router_bwd_dx(unsigned grid, cudaStream_t stream, const T* wg, const float* probs, const float* topk_w,
                          const int32_t* topk_idx, const float* dtw, const float* dprobs, const float* dlogits,
                          const float* lse, const float* cnt, const float* g_losses, long long Tn, int N, int B, int D, int E,
                          int K, int rbf, float* dl, T* dx) {
  if (E <= kMaxE)
    router_bwd_dx_kernel<T><<<grid, kWarpsPerBlock * 32, 0, stream>>>(wg, probs, topk_w, topk_idx, dtw, dprobs, dlogits, lse,
                                                                       cnt, g_losses, Tn, N, B, D, E, K, rbf, dl, dx);
  else if (E <= 128)
    router_bwd_dx_wide_kernel<T, 4><<<grid, kWarpsPerBlock * 32, 0, stream>>>(wg, probs, topk_w, topk_idx, dtw, dprobs, dlogits,
                                                                               lse, cnt, g_losses, Tn, N, B, D, E, K, rbf, dl, dx);
  else
    router_bwd_dx_wide_kernel<T, 8><<<grid, kWarpsPerBlock * 32, 0, stream>>>(wg, probs, topk_w, topk_idx, dtw, dprobs, dlogits,
                                                                               lse, cnt, g_losses, Tn, N, B, D, E, K, rbf, dl, dx);
}

}  // namespace
}  // namespace csmoe

using namespace csmoe;

extern "C" int csmoe_router_fwd(const void* x, const void* wg, int32_t x_dtype, int64_t T, int32_t D, int32_t E,
                                int32_t K, int32_t renorm_dtype, void* logits, float* probs, float* topk_w,
                                int32_t* topk_idx, void* stream_) {
  CSMOE_CHECK_ARG(x && wg && logits && probs && topk_w && topk_idx, "csmoe_router_fwd: NULL pointer");
  CSMOE_CHECK_ARG(E >= 1 && E <= kMaxEWide, "csmoe_router_fwd: E must be in [1, %d], got %d", kMaxEWide, E);
  CSMOE_CHECK_ARG(K >= 1 && K <= kMaxK && K <= E, "csmoe_router_fwd: K must be in [1, min(E, %d)], got %d", kMaxK, K);
  CSMOE_CHECK_ARG(D > 0 && D % 8 == 0, "csmoe_router_fwd: D must be a positive multiple of 8");
  CSMOE_CHECK_ARG(T >= 0, "csmoe_router_fwd: T must be >= 0");
  CSMOE_CHECK_ARG(renorm_dtype == CSMOE_BF16 || renorm_dtype == CSMOE_F32, "csmoe_router_fwd: bad renorm_dtype %d", renorm_dtype);
  if (T == 0) return CSMOE_OK;
  const unsigned grid = static_cast<unsigned>((T + kWarpsPerBlock - 1) / kWarpsPerBlock);
  cudaStream_t stream = as_stream(stream_);
  const int rbf = renorm_dtype == CSMOE_BF16;
  if (x_dtype == CSMOE_BF16) {
    using T_ = __nv_bfloat16;
    launch_router_fwd<T_>(grid, stream, static_cast<const T_*>(x), static_cast<const T_*>(wg), T, D, E, K, rbf,
                          static_cast<T_*>(logits), probs, topk_w, topk_idx);
  } else if (x_dtype == CSMOE_F32) {
    launch_router_fwd<float>(grid, stream, static_cast<const float*>(x), static_cast<const float*>(wg), T, D, E, K, rbf,
                             static_cast<float*>(logits), probs, topk_w, topk_idx);
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_router_fwd: unsupported dtype %d", x_dtype);
  }
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_router_from_logits(const void* logits, int32_t dtype, int64_t T, int32_t E, int32_t K,
                                        int32_t renorm_dtype, float* probs, float* topk_w, int32_t* topk_idx, void* stream_) {
  CSMOE_CHECK_ARG(logits && probs && topk_w && topk_idx, "csmoe_router_from_logits: NULL pointer");
  CSMOE_CHECK_ARG(E >= 1 && E <= kMaxEWide && K >= 1 && K <= kMaxK && K <= E && T >= 0, "csmoe_router_from_logits: bad sizes");
  CSMOE_CHECK_ARG(renorm_dtype == CSMOE_BF16 || renorm_dtype == CSMOE_F32, "csmoe_router_from_logits: bad renorm_dtype %d",
                  renorm_dtype);
  if (T == 0) return CSMOE_OK;
  const unsigned grid = static_cast<unsigned>((T + kWarpsPerBlock - 1) / kWarpsPerBlock);
  cudaStream_t stream = as_stream(stream_);
  const int rbf = renorm_dtype == CSMOE_BF16;
  if (dtype == CSMOE_BF16) {
    launch_router_from_logits<__nv_bfloat16>(grid, stream, static_cast<const __nv_bfloat16*>(logits), T, E, K, rbf, probs, topk_w,
                                             topk_idx);
  } else if (dtype == CSMOE_F32) {
    launch_router_from_logits<float>(grid, stream, static_cast<const float*>(logits), T, E, K, rbf, probs, topk_w, topk_idx);
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_router_from_logits: unsupported dtype %d", dtype);
  }
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_topk_renorm(const float* scores, int64_t T, int32_t E, int32_t K, int32_t mode, int32_t round_dtype,
                                 float* topk_w, int32_t* topk_idx, void* stream_) {
  CSMOE_CHECK_ARG(scores && topk_w && topk_idx, "csmoe_topk_renorm: NULL pointer");
  CSMOE_CHECK_ARG(E >= 1 && E <= kMaxEWide, "csmoe_topk_renorm: E must be in [1, %d]", kMaxEWide);
  CSMOE_CHECK_ARG(K >= 1 && K <= kMaxK && K <= E, "csmoe_topk_renorm: K must be in [1, min(E, %d)]", kMaxK);
  if (T == 0) return CSMOE_OK;
  const unsigned grid = static_cast<unsigned>((T + kWarpsPerBlock - 1) / kWarpsPerBlock);
  if (E > 128)
    topk_renorm_wide_kernel<8><<<grid, kWarpsPerBlock * 32, 0, as_stream(stream_)>>>(scores, T, E, K, mode, round_dtype, topk_w, topk_idx);
  else if (E > kMaxE)
    topk_renorm_wide_kernel<4><<<grid, kWarpsPerBlock * 32, 0, as_stream(stream_)>>>(scores, T, E, K, mode, round_dtype, topk_w, topk_idx);
  else
    topk_renorm_kernel<<<grid, kWarpsPerBlock * 32, 0, as_stream(stream_)>>>(scores, T, E, K, mode, round_dtype, topk_w, topk_idx);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int64_t csmoe_router_aux_workspace_bytes(int64_t B, int64_t N, int32_t E) {
  if (B <= 0 || N <= 0 || E <= 0) return -1;
  const int64_t chunks = (N + kAuxChunk - 1) / kAuxChunk;
  return B * chunks * (2 * E + 1) * static_cast<int64_t>(sizeof(float));
}

extern "C" int csmoe_router_aux_fwd(const void* logits, int32_t dtype, const float* probs, const int32_t* topk_idx,
                                    int64_t B, int64_t N, int32_t E, int32_t K, float* psum, float* cnt, float* lse,
                                    float* losses, void* workspace, void* stream_) {
  CSMOE_CHECK_ARG(logits && probs && topk_idx && psum && cnt && losses && workspace, "csmoe_router_aux_fwd: NULL pointer");
  CSMOE_CHECK_ARG(E >= 1 && E <= kMaxEWide && K >= 1 && B >= 1 && N >= 1, "csmoe_router_aux_fwd: bad sizes");
  cudaStream_t stream = as_stream(stream_);
  const int chunks = static_cast<int>((N + kAuxChunk - 1) / kAuxChunk);
  const unsigned grid = static_cast<unsigned>(B * chunks);
  const size_t smem = kWarpsPerBlock * (2 * E + 1) * sizeof(float);
  float* partial = static_cast<float*>(workspace);
  if (dtype == CSMOE_BF16) {
    launch_router_aux_stage1<__nv_bfloat16>(grid, smem, stream, static_cast<const __nv_bfloat16*>(logits), probs, topk_idx,
                                            static_cast<int>(N), E, K, chunks, partial, lse);
  } else if (dtype == CSMOE_F32) {
    launch_router_aux_stage1<float>(grid, smem, stream, static_cast<const float*>(logits), probs, topk_idx, static_cast<int>(N),
                                    E, K, chunks, partial, lse);
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_router_aux_fwd: unsupported dtype %d", dtype);
  }
  CSMOE_CHECK_LAUNCH();
  router_aux_stage2<<<1, 256, 0, stream>>>(partial, static_cast<int>(B), static_cast<int>(N), E, chunks, psum, cnt, losses);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int64_t csmoe_router_bwd_workspace_bytes(int64_t T, int32_t D, int32_t E) {
  if (T < 0 || D <= 0 || E <= 0) return -1;
  const int64_t chunks = (T + kDwChunk - 1) / kDwChunk;
  return (chunks > 0 ? chunks : 1) * static_cast<int64_t>(E) * D * static_cast<int64_t>(sizeof(float));
}

extern "C" int csmoe_router_bwd(const void* x, const void* wg, int32_t x_dtype, const float* probs, const float* topk_w,
                                const int32_t* topk_idx, const float* dtw, const float* dprobs, const float* dlogits,
                                const float* lse, const float* cnt, const float* g_losses, int64_t B, int64_t N, int32_t D,
                                int32_t E, int32_t K, int32_t renorm_dtype, float* dl, void* dx, void* dwg, int32_t wg_dtype,
                                void* workspace, void* stream_) {
  CSMOE_CHECK_ARG(x && wg && probs && topk_w && topk_idx && dl, "csmoe_router_bwd: NULL pointer");
  CSMOE_CHECK_ARG(E >= 1 && E <= kMaxEWide && K >= 1 && K <= kMaxK && D > 0 && D % 8 == 0 && B >= 1 && N >= 1,
                  "csmoe_router_bwd: bad sizes");
  CSMOE_CHECK_ARG(renorm_dtype == CSMOE_BF16 || renorm_dtype == CSMOE_F32, "csmoe_router_bwd: bad renorm_dtype %d", renorm_dtype);
  CSMOE_CHECK_ARG(dwg == nullptr || workspace != nullptr, "csmoe_router_bwd: dwg needs a workspace");
  cudaStream_t stream = as_stream(stream_);
  const long long T = B * N;
  const unsigned grid = static_cast<unsigned>((T + kWarpsPerBlock - 1) / kWarpsPerBlock);
  const int chunks = static_cast<int>((T + kDwChunk - 1) / kDwChunk);
  dim3 g1((D / 8 + 127) / 128, chunks, (E + kDwExperts - 1) / kDwExperts);
  const long long ED = static_cast<long long>(E) * D;
  const unsigned g2 = static_cast<unsigned>((ED / 8 + 7) / 8);
  float* partial = static_cast<float*>(workspace);
  const int rbf = renorm_dtype == CSMOE_BF16;
  if (x_dtype == CSMOE_BF16) {
    using T_ = __nv_bfloat16;
    launch_router_bwd_dx<T_>(grid, stream, static_cast<const T_*>(wg), probs, topk_w, topk_idx, dtw, dprobs, dlogits, lse, cnt,
                             g_losses, T, static_cast<int>(N), static_cast<int>(B), D, E, K, rbf, dl, static_cast<T_*>(dx));
    CSMOE_CHECK_LAUNCH();
    if (dwg != nullptr) {
      router_bwd_dw_stage1<T_><<<g1, 128, 0, stream>>>(static_cast<const T_*>(x), dl, T, D, E, partial);
      CSMOE_CHECK_LAUNCH();
      if (wg_dtype == CSMOE_BF16)
        router_bwd_dw_stage2<__nv_bfloat16><<<g2, 256, 0, stream>>>(partial, chunks, ED, static_cast<__nv_bfloat16*>(dwg));
      else
        router_bwd_dw_stage2<float><<<g2, 256, 0, stream>>>(partial, chunks, ED, static_cast<float*>(dwg));
      CSMOE_CHECK_LAUNCH();
    }
  } else if (x_dtype == CSMOE_F32) {
    launch_router_bwd_dx<float>(grid, stream, static_cast<const float*>(wg), probs, topk_w, topk_idx, dtw, dprobs, dlogits, lse,
                                cnt, g_losses, T, static_cast<int>(N), static_cast<int>(B), D, E, K, rbf, dl,
                                static_cast<float*>(dx));
    CSMOE_CHECK_LAUNCH();
    if (dwg != nullptr) {
      router_bwd_dw_stage1<float><<<g1, 128, 0, stream>>>(static_cast<const float*>(x), dl, T, D, E, partial);
      CSMOE_CHECK_LAUNCH();
      CSMOE_CHECK_ARG(wg_dtype == CSMOE_F32, "csmoe_router_bwd: fp32 activations need fp32 gate weights");
      router_bwd_dw_stage2<float><<<g2, 256, 0, stream>>>(partial, chunks, ED, static_cast<float*>(dwg));
      CSMOE_CHECK_LAUNCH();
    }
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_router_bwd: unsupported dtype %d", x_dtype);
  }
  return CSMOE_OK;
}
