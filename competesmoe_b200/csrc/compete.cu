// Competition step: neural-response score  aff[t,e] = mean_d softplus(y_e[t,d])  and its backward.
// Replaces `torch.mean(F.softplus(out_i), dim=-1)` (moe_model/model/moe/competesmoe.py:240-243) and
// `torch.mean(F.softplus(expert_outputs), dim=-1)` (moe_pretrain_model/layers/moe/competesmoe.py:399-403).
// One warp reduces one (expert, token) row of the dense expert-output buffer y[E, t_pad, D]; HBM-bound.
#include "common.h"

namespace csmoe {
namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_sp(float x) { return x > 20.f ? 1.f : 1.f / (1.f + expf(-x)); }

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
    for (int c = lane * 8; c < D; c += 256) {
      float v[8];
      load8(row + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float sp = softplus_f(v[j]);
        s += round_bf16 ? bf16_round(sp) : sp;
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
