// Token permutation and combine: coalesced 16-byte gathers / scatters between the token-major layout [T, D] and the
// padded expert-major row space [row_cap, D].  HBM-bound; one warp moves one row, 16 B per lane per access.
// Replaces x[batch_idx, token_idx] / results[batch_idx, token_idx] += w * out of compute_moe
// (moe_model/model/moe/moe.py:199-204) and the gathered loads / out_index stores + reduction bmm of the CVMM op
// (moe_pretrain_model/layers/cvmm.py:114-118,160-167,481-483,538-545).
#include "common.h"

namespace csmoe {
namespace {

constexpr int kWarps = 8;
constexpr int kMaxK = 8;

// dst[row] = scale * src[slot / K]   (zero for padding rows); one warp per (row, block of 1024 columns)
constexpr int kGatherCols = 1024;

template <typename T>
__global__ void __launch_bounds__(kWarps * 32)
gather_rows_kernel(const T* __restrict__ src, int D, int K, const int32_t* __restrict__ row_to_slot, long long row_cap,
                   const float* __restrict__ slot_w, T* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int nblk = (D + kGatherCols - 1) / kGatherCols;
  const long long units = row_cap * nblk;
  for (long long u = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5); u < units;
       u += static_cast<long long>(gridDim.x) * kWarps) {
    const long long row = u / nblk;
    const int c0 = static_cast<int>(u % nblk) * kGatherCols + lane * 8;
    const int slot = row_to_slot[row];
    T* d = dst + row * D;
    if (slot < 0) {
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (c0 + i * 256 < D) store8(d + c0 + i * 256, z);
      continue;
    }
    const T* s = src + static_cast<long long>(slot / K) * D;
    // four 16-byte loads per lane in flight before the first store
    typename Raw8<T>::type raw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (c0 + i * 256 < D) raw[i] = load8_raw(s + c0 + i * 256);
    const float w = slot_w != nullptr ? slot_w[slot] : 1.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (c0 + i * 256 < D) {
        float v[8];
        unpack8(raw[i], v);
        if (slot_w != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] *= w;
        }
        store8(d + c0 + i * 256, v);
      }
    }
  }
}

// The three kernels below are templated on KT = K rounded up to 2 / 4 / 8 (registers and predicated work scale with it:
// with the runtime K <= 8 form the K = 2 shapes ran at 84 registers and 2 CTAs / SM) and, where the work is elementwise
// over columns, one warp owns (token, block of 1024 columns) so that few-token / wide-row shapes (T = 4096, D = 3072)
// still fill the machine.
constexpr int kColBlock = 1024;

// out[t] = sum_k w[t,k] * y[row(t,k)], slots visited in ascending expert id (ties: ascending k).
// flags bit0: round the running sum to T after every term (moe.py:204 accumulates in the output dtype);
// flags bit1: round w to T before use (cvmm.py:483 `reduction_weight.type_as(res)`).
template <typename T, int KT>
__global__ void __launch_bounds__(kWarps * 32)
combine_fwd_kernel(const T* __restrict__ y, long long Tn, int D, int K, const int32_t* __restrict__ slot_to_row,
                   const int32_t* __restrict__ sel, const float* __restrict__ w, int flags, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int nblk = (D + kColBlock - 1) / kColBlock;
  const long long units = Tn * nblk;
  for (long long u = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5); u < units;
       u += static_cast<long long>(gridDim.x) * kWarps) {
    const long long t = u / nblk;
    const int c_lo = static_cast<int>(u % nblk) * kColBlock, c_hi = min(D, c_lo + kColBlock);
    int order[KT];
    int key[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      order[k] = k;
      key[k] = k < K ? sel[t * K + k] : 0x7fffffff;
    }
    // insertion sort (stable) of at most KT keys; lane-uniform
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
      store8(out + t * D + c, acc);
    }
  }
}

// dw[t,k] = <dout[t], y[row(t,k)]>
template <typename T, int KT>
__global__ void __launch_bounds__(kWarps * 32)
combine_bwd_w_kernel(const T* __restrict__ y, const T* __restrict__ dout, long long Tn, int D, int K,
                     const int32_t* __restrict__ slot_to_row, float* __restrict__ dw) {
  const int lane = threadIdx.x & 31;
  for (long long t = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5); t < Tn;
       t += static_cast<long long>(gridDim.x) * kWarps) {
    float acc[KT];
    int rows[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      acc[k] = 0.f;
      rows[k] = k < K ? slot_to_row[t * K + k] : 0;
    }
    for (int c = lane * 8; c < D; c += 256) {
      float g[8];
      typename Raw8<T>::type raw[KT];
      load8(dout + t * D + c, g);
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (k < K) raw[k] = load8_raw(y + static_cast<long long>(rows[k]) * D + c);
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        if (k < K) {
          float v[8];
          unpack8(raw[k], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[k] = fmaf(g[j], v[j], acc[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      if (k < K) {
        float s = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) dw[t * K + k] = s;
      }
    }
  }
}

// dx[t] (+)= sum_k g[row(t,k)]   (fp32 sum in k order, one rounding)
template <typename T, int KT>
__global__ void __launch_bounds__(kWarps * 32)
scatter_reduce_kernel(const T* __restrict__ g, long long Tn, int D, int K, const int32_t* __restrict__ slot_to_row,
                      int accumulate, T* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const int nblk = (D + kColBlock - 1) / kColBlock;
  const long long units = Tn * nblk;
  for (long long u = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5); u < units;
       u += static_cast<long long>(gridDim.x) * kWarps) {
    const long long t = u / nblk;
    const int c_lo = static_cast<int>(u % nblk) * kColBlock, c_hi = min(D, c_lo + kColBlock);
    int rows[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) rows[k] = k < K ? slot_to_row[t * K + k] : 0;
    for (int c = c_lo + lane * 8; c < c_hi; c += 256) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      typename Raw8<T>::type raw[KT];
      if (accumulate) load8(dx + t * D + c, acc);
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (k < K) raw[k] = load8_raw(g + static_cast<long long>(rows[k]) * D + c);
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        if (k < K) {
          float v[8];
          unpack8(raw[k], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += v[j];
        }
      }
      store8(dx + t * D + c, acc);
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

#define DISPATCH_KT(K, ...)                                  \
  if ((K) <= 2) { constexpr int KT = 2; __VA_ARGS__; }       \
  else if ((K) <= 4) { constexpr int KT = 4; __VA_ARGS__; }   \
  else { constexpr int KT = 8; __VA_ARGS__; }

#define DISPATCH_DTYPE(dtype, ...)                          \
  if ((dtype) == CSMOE_BF16) {                              \
    using T = __nv_bfloat16;                                \
    __VA_ARGS__;                                            \
  } else if ((dtype) == CSMOE_F32) {                        \
    using T = float;                                        \
    __VA_ARGS__;                                            \
  } else {                                                  \
    CSMOE_CHECK_ARG(false, "unsupported dtype %d", (dtype)); \
  }

extern "C" int csmoe_gather_rows(const void* src, int32_t dtype, int64_t T_, int32_t D, int32_t K,
                                 const int32_t* row_to_slot, int64_t row_cap, const float* slot_w, void* dst,
                                 void* stream_) {
  CSMOE_CHECK_ARG(src && row_to_slot && dst, "csmoe_gather_rows: NULL pointer");
  CSMOE_CHECK_ARG(D > 0 && D % 8 == 0 && K >= 1, "csmoe_gather_rows: D must be a multiple of 8, K >= 1");
  (void)T_;
  if (row_cap == 0) return CSMOE_OK;
  cudaStream_t stream = as_stream(stream_);
  DISPATCH_DTYPE(dtype, (gather_rows_kernel<T><<<row_grid(row_cap * ((D + kGatherCols - 1) / kGatherCols)), kWarps * 32, 0, stream>>>(
                            static_cast<const T*>(src), D, K, row_to_slot, row_cap, slot_w, static_cast<T*>(dst))));
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_combine_fwd(const void* y, int32_t dtype, int64_t T_, int32_t D, int32_t K,
                                 const int32_t* slot_to_row, const int32_t* sel, const float* w, int32_t flags,
                                 void* out, void* stream_) {
  CSMOE_CHECK_ARG(y && slot_to_row && sel && w && out, "csmoe_combine_fwd: NULL pointer");
  CSMOE_CHECK_ARG(D > 0 && D % 8 == 0 && K >= 1 && K <= kMaxK, "csmoe_combine_fwd: D %% 8 == 0 and 1 <= K <= %d", kMaxK);
  if (T_ == 0) return CSMOE_OK;
  cudaStream_t stream = as_stream(stream_);
  const long long units = T_ * ((D + kColBlock - 1) / kColBlock);
  DISPATCH_DTYPE(dtype, DISPATCH_KT(K, (combine_fwd_kernel<T, KT><<<row_grid(units), kWarps * 32, 0, stream>>>(
                                          static_cast<const T*>(y), T_, D, K, slot_to_row, sel, w, flags, static_cast<T*>(out)))));
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_combine_bwd_w(const void* y, const void* dout, int32_t dtype, int64_t T_, int32_t D, int32_t K,
                                   const int32_t* slot_to_row, float* dw, void* stream_) {
  CSMOE_CHECK_ARG(y && dout && slot_to_row && dw, "csmoe_combine_bwd_w: NULL pointer");
  CSMOE_CHECK_ARG(D > 0 && D % 8 == 0 && K >= 1 && K <= kMaxK, "csmoe_combine_bwd_w: D %% 8 == 0 and 1 <= K <= %d", kMaxK);
  if (T_ == 0) return CSMOE_OK;
  cudaStream_t stream = as_stream(stream_);
  DISPATCH_DTYPE(dtype, DISPATCH_KT(K, (combine_bwd_w_kernel<T, KT><<<row_grid(T_), kWarps * 32, 0, stream>>>(
                                          static_cast<const T*>(y), static_cast<const T*>(dout), T_, D, K, slot_to_row, dw))));
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_scatter_reduce(const void* g, int32_t dtype, int64_t T_, int32_t D, int32_t K,
                                    const int32_t* slot_to_row, int32_t accumulate, void* dx, void* stream_) {
  CSMOE_CHECK_ARG(g && slot_to_row && dx, "csmoe_scatter_reduce: NULL pointer");
  CSMOE_CHECK_ARG(D > 0 && D % 8 == 0 && K >= 1 && K <= kMaxK, "csmoe_scatter_reduce: D %% 8 == 0 and 1 <= K <= %d", kMaxK);
  if (T_ == 0) return CSMOE_OK;
  cudaStream_t stream = as_stream(stream_);
  const long long units = T_ * ((D + kColBlock - 1) / kColBlock);
  DISPATCH_DTYPE(dtype, DISPATCH_KT(K, (scatter_reduce_kernel<T, KT><<<row_grid(units), kWarps * 32, 0, stream>>>(
                                          static_cast<const T*>(g), T_, D, K, slot_to_row, accumulate, static_cast<T*>(dx)))));
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}
