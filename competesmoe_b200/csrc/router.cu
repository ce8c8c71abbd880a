// Router: skinny gate GEMM fused with fp32 softmax, warp-shuffle top-k and renormalisation (one warp per token).
// Replaces router_policy + topk_expert (moe_model/model/moe/competesmoe.py:301-320, moe.py:113-132;
// moe_pretrain_model/layers/moe/competesmoe.py:465-490, moe.py:373-393).
//
// Rounding points follow the reference: logits are accumulated in fp32 and rounded to the activation dtype (the
// nn.Linear / F.linear output), softmax runs in fp32 on the rounded logits, the top-k weights stay fp32 and are divided
// by their sum rounded to the activation dtype (`.to(x.dtype)` on the denominator only).
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

// Warp-wide selection of the K largest of up to 64 values (lane l holds experts l and l+32).
// Emits (value, index) pairs in descending order to lane-uniform arrays.
__device__ __forceinline__ void warp_topk(float v0, float v1, int lane, int E, int K, float (&out_v)[kMaxK],
                                          int (&out_i)[kMaxK]) {
  const float NEG = -INFINITY;
  if (lane >= E) v0 = NEG;
  if (lane + 32 >= E) v1 = NEG;
  bool t0 = lane >= E, t1 = lane + 32 >= E;  // taken / invalid
#pragma unroll
  for (int k = 0; k < kMaxK; ++k) {
    if (k >= K) break;
    float bv;
    int bi;
    // local best (lower index wins ties): candidate 0 has the lower index
    if (!t0 && (t1 || v0 >= v1)) {
      bv = v0;
      bi = lane;
    } else if (!t1) {
      bv = v1;
      bi = lane + 32;
    } else {
      bv = NEG;
      bi = 0x7fffffff;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    out_v[k] = bv;
    out_i[k] = bi;
    if (bi == lane) t0 = true;
    if (bi == lane + 32) t1 = true;
  }
}

template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
router_fwd_kernel(const T* __restrict__ x, const T* __restrict__ wg, long long Tn, int D, int E, int K,
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
    s = round_as(s, static_cast<const T*>(nullptr));
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

}  // namespace
}  // namespace csmoe

using namespace csmoe;

extern "C" int csmoe_router_fwd(const void* x, const void* wg, int32_t x_dtype, int64_t T, int32_t D, int32_t E,
                                int32_t K, void* logits, float* probs, float* topk_w, int32_t* topk_idx, void* stream_) {
  CSMOE_CHECK_ARG(x && wg && logits && probs && topk_w && topk_idx, "csmoe_router_fwd: NULL pointer");
  CSMOE_CHECK_ARG(E >= 1 && E <= kMaxE, "csmoe_router_fwd: E must be in [1, %d], got %d", kMaxE, E);
  CSMOE_CHECK_ARG(K >= 1 && K <= kMaxK && K <= E, "csmoe_router_fwd: K must be in [1, min(E, %d)], got %d", kMaxK, K);
  CSMOE_CHECK_ARG(D > 0 && D % 8 == 0, "csmoe_router_fwd: D must be a positive multiple of 8");
  CSMOE_CHECK_ARG(T >= 0, "csmoe_router_fwd: T must be >= 0");
  if (T == 0) return CSMOE_OK;
  const unsigned grid = static_cast<unsigned>((T + kWarpsPerBlock - 1) / kWarpsPerBlock);
  cudaStream_t stream = as_stream(stream_);
  if (x_dtype == CSMOE_BF16) {
    router_fwd_kernel<__nv_bfloat16><<<grid, kWarpsPerBlock * 32, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(wg), T, D, E, K,
        static_cast<__nv_bfloat16*>(logits), probs, topk_w, topk_idx);
  } else if (x_dtype == CSMOE_F32) {
    router_fwd_kernel<float><<<grid, kWarpsPerBlock * 32, 0, stream>>>(static_cast<const float*>(x),
                                                                       static_cast<const float*>(wg), T, D, E, K,
                                                                       static_cast<float*>(logits), probs, topk_w, topk_idx);
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_router_fwd: unsupported dtype %d", x_dtype);
  }
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_topk_renorm(const float* scores, int64_t T, int32_t E, int32_t K, int32_t mode, int32_t round_dtype,
                                 float* topk_w, int32_t* topk_idx, void* stream_) {
  CSMOE_CHECK_ARG(scores && topk_w && topk_idx, "csmoe_topk_renorm: NULL pointer");
  CSMOE_CHECK_ARG(E >= 1 && E <= kMaxE, "csmoe_topk_renorm: E must be in [1, %d]", kMaxE);
  CSMOE_CHECK_ARG(K >= 1 && K <= kMaxK && K <= E, "csmoe_topk_renorm: K must be in [1, min(E, %d)]", kMaxK);
  if (T == 0) return CSMOE_OK;
  const unsigned grid = static_cast<unsigned>((T + kWarpsPerBlock - 1) / kWarpsPerBlock);
  topk_renorm_kernel<<<grid, kWarpsPerBlock * 32, 0, as_stream(stream_)>>>(scores, T, E, K, mode, round_dtype, topk_w,
                                                                           topk_idx);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}
