// Shared host-side helpers for the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "csmoe.h"

namespace csmoe {

void set_error(const char* fmt, ...);
int num_sms();

#define CSMOE_CHECK_ARG(cond, ...)       \
  do {                                   \
    if (!(cond)) {                       \
      ::csmoe::set_error(__VA_ARGS__);   \
      return CSMOE_ERR_ARG;              \
    }                                    \
  } while (0)

#define CSMOE_CHECK_CUDA(expr)                                                               \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::csmoe::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CSMOE_ERR_CUDA;                                                                 \
    }                                                                                        \
  } while (0)

#define CSMOE_CHECK_LAUNCH() CSMOE_CHECK_CUDA(cudaGetLastError())

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- device helpers shared by the bandwidth-bound kernels
template <typename T>
struct Vec8;  // 8 elements = one 16-byte (bf16) or two 16-byte (fp32) accesses

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
// Raw 8-element loads (kept packed while several are in flight: 4 registers per bf16 vector instead of 8 floats).
struct Raw8bf { uint4 u; };
struct Raw8f { float4 a, b; };
__device__ __forceinline__ Raw8bf load8_raw(const __nv_bfloat16* p) { return Raw8bf{*reinterpret_cast<const uint4*>(p)}; }
__device__ __forceinline__ Raw8f load8_raw(const float* p) {
  return Raw8f{*reinterpret_cast<const float4*>(p), *reinterpret_cast<const float4*>(p + 4)};
}
__device__ __forceinline__ void unpack8(const Raw8bf& r, float (&f)[8]) {
  const uint32_t w[4] = {r.u.x, r.u.y, r.u.z, r.u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void unpack8(const Raw8f& r, float (&f)[8]) {
  f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w;
  f[4] = r.b.x; f[5] = r.b.y; f[6] = r.b.z; f[7] = r.b.w;
}
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> { using type = Raw8bf; };
template <> struct Raw8<float> { using type = Raw8f; };

__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ float round_as(float x, const __nv_bfloat16*) { return bf16_round(x); }
__device__ __forceinline__ float round_as(float x, const float*) { return x; }

// ---- activations (fp32 math on values already rounded to the storage dtype, like the reference's eager ops)
// `fast` (bf16 outputs only): GELU-tanh uses the hardware tanh.approx.f32 (max rel. error 2^-11, four times finer than a
// bf16 ulp) instead of tanhf -- the precise version made the GELU epilogue of short-k GEMMs (SigLIP fc1: k = 1152) and
// the stand-alone activation backward instruction bound.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float act_apply(float z, int act, bool fast = false) {
  switch (act) {
    case CSMOE_ACT_RELU:
      return z > 0.f ? z : 0.f;
    case CSMOE_ACT_GELU:
      return 0.5f * z * (1.f + erff(z * 0.70710678118654752440f));
    case CSMOE_ACT_GELU_TANH: {
      const float k0 = 0.79788456080286535588f, k1 = 0.044715f;
      const float u = k0 * (z + k1 * z * z * z);
      return 0.5f * z * (1.f + (fast ? tanh_approx(u) : tanhf(u)));
    }
    case CSMOE_ACT_SILU:
      return fast ? __fdividef(z, 1.f + __expf(-z)) : z / (1.f + __expf(-z));
    default:
      return z;
  }
}
__device__ __forceinline__ float act_grad(float z, int act, bool fast = false) {
  switch (act) {
    case CSMOE_ACT_RELU:
      return z > 0.f ? 1.f : 0.f;
    case CSMOE_ACT_GELU: {
      const float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * __expf(-0.5f * z * z);
      return cdf + z * pdf;
    }
    case CSMOE_ACT_GELU_TANH: {
      const float k0 = 0.79788456080286535588f, k1 = 0.044715f;
      const float u = k0 * (z + k1 * z * z * z);
      const float t = fast ? tanh_approx(u) : tanhf(u);
      const float du = k0 * (1.f + 3.f * k1 * z * z);
      return 0.5f * (1.f + t) + 0.5f * z * (1.f - t * t) * du;
    }
    case CSMOE_ACT_SILU: {
      const float s = fast ? __fdividef(1.f, 1.f + __expf(-z)) : 1.f / (1.f + __expf(-z));
      return s * (1.f + z * (1.f - s));
    }
    default:
      return 1.f;
  }
}

// Whole-vector versions with the activation switch hoisted out of the element loop (one branch per N elements; inside a
// GEMM epilogue the per-element switch kept every case's code on the hot path and made the activation epilogues 2.5x
// slower than the plain one).
template <int N>
__device__ __forceinline__ void act_apply_vec(float (&z)[N], int act, bool fast) {
  switch (act) {
    case CSMOE_ACT_RELU:
#pragma unroll
      for (int i = 0; i < N; ++i) z[i] = fmaxf(z[i], 0.f);
      break;
    case CSMOE_ACT_GELU:
#pragma unroll
      for (int i = 0; i < N; ++i) z[i] = act_apply(z[i], CSMOE_ACT_GELU, fast);
      break;
    case CSMOE_ACT_GELU_TANH:
      if (fast) {
#pragma unroll
        for (int i = 0; i < N; ++i) z[i] = act_apply(z[i], CSMOE_ACT_GELU_TANH, true);
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) z[i] = act_apply(z[i], CSMOE_ACT_GELU_TANH, false);
      }
      break;
    case CSMOE_ACT_SILU:
#pragma unroll
      for (int i = 0; i < N; ++i) z[i] = act_apply(z[i], CSMOE_ACT_SILU, fast);
      break;
    default:
      break;
  }
}

// g[i] *= act'(z[i])
template <int N>
__device__ __forceinline__ void act_grad_vec(float (&g)[N], const float (&z)[N], int act, bool fast) {
  switch (act) {
    case CSMOE_ACT_RELU:
#pragma unroll
      for (int i = 0; i < N; ++i) g[i] = z[i] > 0.f ? g[i] : 0.f;
      break;
    case CSMOE_ACT_GELU:
#pragma unroll
      for (int i = 0; i < N; ++i) g[i] *= act_grad(z[i], CSMOE_ACT_GELU, fast);
      break;
    case CSMOE_ACT_GELU_TANH:
      if (fast) {
#pragma unroll
        for (int i = 0; i < N; ++i) g[i] *= act_grad(z[i], CSMOE_ACT_GELU_TANH, true);
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) g[i] *= act_grad(z[i], CSMOE_ACT_GELU_TANH, false);
      }
      break;
    case CSMOE_ACT_SILU:
#pragma unroll
      for (int i = 0; i < N; ++i) g[i] *= act_grad(z[i], CSMOE_ACT_SILU, fast);
      break;
    default:
      break;
  }
}

}  // namespace csmoe
