// Expert parallelism over NVLink 5 / NVSwitch peer memory (stage 6 of the path).
//
// One process per GPU.  Every rank owns E / P consecutive experts.  Each rank maps the other ranks' exchange buffers
// into its address space once (CUDA IPC), after which dispatch and return are plain kernels that store straight into
// peer HBM -- there is no staging buffer, no host-side split computation and no host synchronisation:
//
//   exchange+plan  every rank publishes its per-expert row counts into all peers' count matrix, a flag barrier makes
//                  the P x E matrix complete everywhere, and each rank derives from it (a) where its own rows start in
//                  every owner's row space and (b) the padded expert-major layout of the rows it will receive.
//   dispatch       permutation fused with the transfer: the row of slot (t, k) goes directly to its place in the
//                  owner's padded expert-major receive buffer (plus one 8-byte tag naming the source slot).
//   return         the grouped GEMM's epilogue stores every output row at `c_rows[row]`, a pointer into the *source*
//                  rank's return buffer (slot order), so the down projection and the combine-side transfer are one
//                  kernel; csmoe_ep_push_rows is the same transfer as a stand-alone kernel.
//   barrier        flag exchange with release/acquire at system scope; the epoch counter lives in device memory, so
//                  the sequence is CUDA-graph capturable.  Spins are bounded and trap.
//
// The reference has no counterpart (it is data-parallel only: moe_pretrain_model/framework/task/simple_task.py:403-413,
// DeepSpeed ZeRO in moe_model/train/train.py:1474-1480).
#include <algorithm>
#include <cstring>

#include "common.h"

namespace csmoe {
namespace {

constexpr int kMaxRanks = CSMOE_EP_MAX_RANKS;
constexpr int kMaxExperts = 1024;
constexpr unsigned long long kBarrierTimeoutNs = 60ull * 1000ull * 1000ull * 1000ull;

struct Peers {
  void* p[kMaxRanks];
};

__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All threads of the (single) CTA call this.  flags.p[r] = rank r's flag array [P]; epoch = this rank's counter.
__device__ __forceinline__ void flag_barrier(const Peers& flags, int* epoch, int rank, int P) {
  __shared__ int s_ep;
  __syncthreads();
  if (threadIdx.x == 0) {
    s_ep = *epoch + 1;
    *epoch = s_ep;
  }
  __threadfence_system();
  __syncthreads();
  const int ep = s_ep;
  if (threadIdx.x < P) {
    st_release_sys(reinterpret_cast<int*>(flags.p[threadIdx.x]) + rank, ep);
    const int* mine = reinterpret_cast<const int*>(flags.p[rank]) + threadIdx.x;
    // a peer died or the ranks disagree on the call sequence: trap after kBarrierTimeoutNs of wall clock (an error on
    // this rank within a minute instead of a job that hangs until someone kills it)
    unsigned long long spins = 0, t0 = 0;
    while (ld_acquire_sys(mine) - ep < 0) {
      if ((++spins & 1023ull) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > kBarrierTimeoutNs) __trap();
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(32) ep_barrier_kernel(Peers flags, int* epoch, int rank, int P) {
  flag_barrier(flags, epoch, rank, P);
}

// Publish counts -> barrier -> plan.  One CTA.
//   counts_all (peer array): [P][E] int32 on every rank.
//   dest_base[E]        row, in the owner's receive space, where THIS rank's rows for expert e start
//   recv_counts[El], recv_pad_offsets[El+1], tile_expert[row_cap/128]  layout of the rows this rank receives
__global__ void __launch_bounds__(256)
ep_exchange_plan_kernel(const int32_t* __restrict__ counts, Peers counts_all, Peers flags, int* epoch, int rank, int P,
                        int E, int row_tile, long long row_cap, int32_t* __restrict__ dest_base,
                        int32_t* __restrict__ recv_counts, int32_t* __restrict__ recv_pad_offsets,
                        int32_t* __restrict__ tile_expert) {
  __shared__ int32_t s_total[kMaxExperts];
  __shared__ int32_t s_pad[kMaxExperts + kMaxRanks];   // padded start of e inside its owner's space
  __shared__ int32_t s_end[kMaxRanks];                 // padded end of each owner's space
  const int El = E / P;
  for (int p = 0; p < P; ++p) {
    int32_t* dst = reinterpret_cast<int32_t*>(counts_all.p[p]) + static_cast<long long>(rank) * E;
    for (int e = threadIdx.x; e < E; e += blockDim.x) dst[e] = counts[e];
  }
  flag_barrier(flags, epoch, rank, P);
  const volatile int32_t* C = reinterpret_cast<const volatile int32_t*>(counts_all.p[rank]);
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    int tot = 0, before = 0;
    for (int s = 0; s < P; ++s) {
      const int v = C[static_cast<long long>(s) * E + e];
      tot += v;
      if (s < rank) before += v;
    }
    s_total[e] = tot;
    dest_base[e] = before;   // completed below with the padded start
  }
  __syncthreads();
  if (threadIdx.x < P) {
    const int o = threadIdx.x;
    int po = 0;
    for (int el = 0; el < El; ++el) {
      s_pad[o * El + el] = po;
      po += (s_total[o * El + el] + row_tile - 1) / row_tile * row_tile;
    }
    s_end[o] = po;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) dest_base[e] += s_pad[e];
  for (int el = threadIdx.x; el < El; el += blockDim.x) {
    recv_counts[el] = s_total[rank * El + el];
    recv_pad_offsets[el] = s_pad[rank * El + el];
  }
  if (threadIdx.x == 0) recv_pad_offsets[El] = s_end[rank];
  const long long n_tiles = row_cap / CSMOE_ROW_TILE;
  for (long long t = threadIdx.x; t < n_tiles; t += blockDim.x) {
    const long long r = t * CSMOE_ROW_TILE;
    int owner = -1;
    if (r < s_end[rank]) {
      // binary search over the local experts' padded starts
      int lo = 0, hi = El - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_pad[rank * El + mid] <= r) lo = mid; else hi = mid - 1;
      }
      owner = lo;
    }
    tile_expert[t] = owner;
  }
}

// Permute + send: the source row of slot j = t*K + k goes to row dest_base[e] + rank_in_expert(j) of its owner.
template <typename T>
__global__ void __launch_bounds__(256)
ep_dispatch_kernel(const T* __restrict__ src, int D, int K, long long n_slots, const int32_t* __restrict__ sel,
                   const int32_t* __restrict__ slot_to_row, const int32_t* __restrict__ pad_offsets,
                   const int32_t* __restrict__ dest_base, int El, const float* __restrict__ slot_w, Peers recv,
                   Peers tags, int rank) {
  const int lane = threadIdx.x & 31;
  for (long long j = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5); j < n_slots;
       j += static_cast<long long>(gridDim.x) * 8) {
    const int e = sel[j];
    const int owner = e / El;
    const long long row = dest_base[e] + (slot_to_row[j] - pad_offsets[e]);
    const T* s = src + (j / K) * D;
    T* d = reinterpret_cast<T*>(recv.p[owner]) + row * D;
    if (slot_w == nullptr) {
      for (int c = lane * 8; c < D; c += 256) {
        float v[8];
        load8(s + c, v);
        store8(d + c, v);
      }
    } else {
      const float w = slot_w[j];
      for (int c = lane * 8; c < D; c += 256) {
        float v[8];
        load8(s + c, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= w;
        store8(d + c, v);
      }
    }
    if (lane == 0 && tags.p[owner] != nullptr)
      reinterpret_cast<long long*>(tags.p[owner])[row] = (static_cast<long long>(rank) << 32) | j;
  }
}

// Receiver side, after the dispatch barrier: c_rows[row] = address of the source slot's row in the source rank's
// return buffer (0 for padding / unused rows), and the padding rows of the receive buffer are zeroed (the wgrad
// GEMM contracts over whole padded segments).
template <typename T>
__global__ void __launch_bounds__(256)
ep_row_ptrs_kernel(const long long* __restrict__ tags, const int32_t* __restrict__ tile_expert,
                   const int32_t* __restrict__ recv_counts, const int32_t* __restrict__ recv_pad_offsets, int El,
                   long long row_cap, Peers ret, long long ret_ld, unsigned long long* __restrict__ c_rows,
                   T* __restrict__ recv, int D) {
  const int lane = threadIdx.x & 31;
  const long long used = recv_pad_offsets[El];
  for (long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5); r < row_cap;
       r += static_cast<long long>(gridDim.x) * 8) {
    bool valid = false;
    if (r < used) {
      const int el = tile_expert[r / CSMOE_ROW_TILE];
      valid = el >= 0 && (r - recv_pad_offsets[el]) < recv_counts[el];
    }
    if (valid) {
      if (lane == 0 && c_rows != nullptr) {
        const long long tag = tags[r];
        const int srcr = static_cast<int>(tag >> 32);
        const long long slot = tag & 0xffffffffll;
        c_rows[r] = reinterpret_cast<unsigned long long>(reinterpret_cast<T*>(ret.p[srcr]) + slot * ret_ld);
      }
    } else {
      if (lane == 0 && c_rows != nullptr) c_rows[r] = 0ull;
      if (recv != nullptr && r < used) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int c = lane * 8; c < D; c += 256) store8(recv + r * D + c, z);
      }
    }
  }
}

// Stand-alone return transfer: dst_rows[r] <- src[r, :]  for every row with a non-null destination.
template <typename T>
__global__ void __launch_bounds__(256)
ep_push_rows_kernel(const T* __restrict__ src, long long ld, int D, long long rows,
                    const unsigned long long* __restrict__ dst_rows) {
  const int lane = threadIdx.x & 31;
  for (long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5); r < rows;
       r += static_cast<long long>(gridDim.x) * 8) {
    T* d = reinterpret_cast<T*>(dst_rows[r]);
    if (d == nullptr) continue;
    for (int c = lane * 8; c < D; c += 256) {
      float v[8];
      load8(src + r * ld + c, v);
      store8(d + c, v);
    }
  }
}

// ---- weight exchange (sigma-MoE shapes: the experts are small, the K-fold expanded token rows are not)
// Every rank keeps its E / P experts' parameters and optimizer state; for compute, every rank holds a bf16 copy of ALL
// experts in a symmetric buffer.  gather_push = cast + all-gather in one pass (each 16-byte vector of the local shard is
// converted once and stored into every rank's copy, peers staggered by rank so that the P writers do not converge on
// one destination); reduce_pull = reduce-scatter: the owner sums its slice of every rank's full-size gradient buffer
// in ascending rank order (deterministic), all P loads of a vector in flight together.
template <typename S, typename T>
__global__ void __launch_bounds__(256)
ep_gather_push_kernel(const S* __restrict__ src, long long n, Peers dst, long long dst_off, int rank, int P) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    float v[8];
    load8(src + i, v);
    for (int q = 0; q < P; ++q) {
      int p = rank + 1 + q;
      if (p >= P) p -= P;
      store8(reinterpret_cast<T*>(dst.p[p]) + dst_off + i, v);
    }
  }
}

// PMAX = compile-time bound on P, U = vectors per thread and iteration: U * (P - 1) remote 16-byte loads are in flight
// per thread (small groups need U > 1 to cover the NVLink latency: one vector per thread measured 410 GB/s at P = 2).
template <typename T, int PMAX, int U>
__global__ void __launch_bounds__(256)
ep_reduce_pull_kernel(Peers src, long long src_off, long long n, T* __restrict__ out, int P) {
  const long long tile = static_cast<long long>(blockDim.x) * 4 * U;
  for (long long base = static_cast<long long>(blockIdx.x) * tile; base < n; base += static_cast<long long>(gridDim.x) * tile) {
    float4 v[U][PMAX];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = base + (static_cast<long long>(u) * blockDim.x + threadIdx.x) * 4;
#pragma unroll
      for (int p = 0; p < PMAX; ++p)
        if (p < P && i < n) v[u][p] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src.p[p]) + src_off + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = base + (static_cast<long long>(u) * blockDim.x + threadIdx.x) * 4;
      if (i >= n) continue;
      float4 a = v[u][0];
#pragma unroll
      for (int p = 1; p < PMAX; ++p)
        if (p < P) { a.x += v[u][p].x; a.y += v[u][p].y; a.z += v[u][p].z; a.w += v[u][p].w; }
      if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(out + i) = a;
      } else {
        __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
        uint2 w;
        w.x = *reinterpret_cast<uint32_t*>(&lo);
        w.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(out + i) = w;
      }
    }
  }
}

template <typename T>
void launch_reduce_pull(const Peers& s, long long src_off, long long n, T* out, int P, cudaStream_t stream) {
  const long long cap = static_cast<long long>(num_sms() > 0 ? num_sms() : 148) * 4;
  auto grid = [&](int U) {
    const long long tile = 256ll * 4 * U;
    return static_cast<unsigned>(std::min<long long>((n + tile - 1) / tile, cap));
  };
  if (P <= 2) ep_reduce_pull_kernel<T, 2, 4><<<grid(4), 256, 0, stream>>>(s, src_off, n, out, P);
  else if (P <= 4) ep_reduce_pull_kernel<T, 4, 2><<<grid(2), 256, 0, stream>>>(s, src_off, n, out, P);
  else if (P <= 8) ep_reduce_pull_kernel<T, 8, 1><<<grid(1), 256, 0, stream>>>(s, src_off, n, out, P);
  else ep_reduce_pull_kernel<T, kMaxRanks, 1><<<grid(1), 256, 0, stream>>>(s, src_off, n, out, P);
}

int fill_peers(Peers& out, const void* const* ptrs, int P, bool allow_null, const char* what) {
  for (int i = 0; i < kMaxRanks; ++i) out.p[i] = nullptr;
  if (ptrs == nullptr) {
    if (allow_null) return CSMOE_OK;
    set_error("%s: peer pointer array is NULL", what);
    return CSMOE_ERR_ARG;
  }
  for (int i = 0; i < P; ++i) {
    if (ptrs[i] == nullptr && !allow_null) {
      set_error("%s: peer pointer %d is NULL", what, i);
      return CSMOE_ERR_ARG;
    }
    out.p[i] = const_cast<void*>(ptrs[i]);
  }
  return CSMOE_OK;
}

inline unsigned warp_grid(long long rows) {
  const long long blocks = (rows + 7) / 8;
  const long long cap = static_cast<long long>(num_sms() > 0 ? num_sms() : 148) * 8;
  return static_cast<unsigned>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace
}  // namespace csmoe

using namespace csmoe;

extern "C" int csmoe_ep_ipc_handle_bytes(void) { return static_cast<int>(sizeof(cudaIpcMemHandle_t)); }

extern "C" int csmoe_ep_alloc(int64_t bytes, void** ptr, void* handle_out) {
  CSMOE_CHECK_ARG(bytes > 0 && ptr != nullptr, "csmoe_ep_alloc: bytes must be positive and ptr non-NULL");
  void* p = nullptr;
  CSMOE_CHECK_CUDA(cudaMalloc(&p, static_cast<size_t>(bytes)));
  cudaError_t e = cudaMemset(p, 0, static_cast<size_t>(bytes));
  if (e == cudaSuccess && handle_out != nullptr)
    e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_out), p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("csmoe_ep_alloc: %s", cudaGetErrorString(e));
    return CSMOE_ERR_CUDA;
  }
  CSMOE_CHECK_CUDA(cudaDeviceSynchronize());
  *ptr = p;
  return CSMOE_OK;
}

extern "C" int csmoe_ep_open(const void* handle, void** ptr) {
  CSMOE_CHECK_ARG(handle != nullptr && ptr != nullptr, "csmoe_ep_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  CSMOE_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return CSMOE_OK;
}

extern "C" int csmoe_ep_close(void* ptr) {
  if (ptr != nullptr) CSMOE_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return CSMOE_OK;
}

extern "C" int csmoe_ep_free(void* ptr) {
  if (ptr != nullptr) CSMOE_CHECK_CUDA(cudaFree(ptr));
  return CSMOE_OK;
}

extern "C" int csmoe_ep_barrier(const void* const* flags, int32_t* epoch, int32_t rank, int32_t P, void* stream_) {
  CSMOE_CHECK_ARG(P >= 1 && P <= kMaxRanks && rank >= 0 && rank < P && epoch != nullptr, "csmoe_ep_barrier: bad rank / P");
  Peers f;
  int rc = fill_peers(f, flags, P, false, "csmoe_ep_barrier");
  if (rc != CSMOE_OK) return rc;
  ep_barrier_kernel<<<1, 32, 0, as_stream(stream_)>>>(f, epoch, rank, P);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_ep_exchange_plan(const int32_t* counts, const void* const* counts_all, const void* const* flags,
                                      int32_t* epoch, int32_t rank, int32_t P, int32_t E, int32_t row_tile,
                                      int64_t row_cap, int32_t* dest_base, int32_t* recv_counts,
                                      int32_t* recv_pad_offsets, int32_t* tile_expert, void* stream_) {
  CSMOE_CHECK_ARG(P >= 1 && P <= kMaxRanks && rank >= 0 && rank < P, "csmoe_ep_exchange_plan: bad rank / P");
  CSMOE_CHECK_ARG(E >= P && E % P == 0 && E <= kMaxExperts, "csmoe_ep_exchange_plan: E (%d) must be a multiple of P (%d), <= %d",
                  E, P, kMaxExperts);
  CSMOE_CHECK_ARG(row_tile == 128 || row_tile == 256, "csmoe_ep_exchange_plan: row_tile must be 128 or 256");
  CSMOE_CHECK_ARG(row_cap > 0 && row_cap % row_tile == 0, "csmoe_ep_exchange_plan: row_cap must be a multiple of row_tile");
  CSMOE_CHECK_ARG(counts && epoch && dest_base && recv_counts && recv_pad_offsets && tile_expert,
                  "csmoe_ep_exchange_plan: NULL pointer");
  Peers ca, f;
  int rc = fill_peers(ca, counts_all, P, false, "csmoe_ep_exchange_plan(counts_all)");
  if (rc != CSMOE_OK) return rc;
  rc = fill_peers(f, flags, P, false, "csmoe_ep_exchange_plan(flags)");
  if (rc != CSMOE_OK) return rc;
  ep_exchange_plan_kernel<<<1, 256, 0, as_stream(stream_)>>>(counts, ca, f, epoch, rank, P, E, row_tile, row_cap, dest_base,
                                                             recv_counts, recv_pad_offsets, tile_expert);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

#define EP_DISPATCH_DTYPE(dtype, ...)                        \
  if ((dtype) == CSMOE_BF16) {                               \
    using T = __nv_bfloat16;                                 \
    __VA_ARGS__;                                             \
  } else if ((dtype) == CSMOE_F32) {                         \
    using T = float;                                         \
    __VA_ARGS__;                                             \
  } else {                                                   \
    CSMOE_CHECK_ARG(false, "unsupported dtype %d", (dtype)); \
  }

extern "C" int csmoe_ep_dispatch(const void* src, int32_t dtype, int32_t D, int32_t K, int64_t n_slots, const int32_t* sel,
                                 const int32_t* slot_to_row, const int32_t* pad_offsets, const int32_t* dest_base,
                                 int32_t experts_per_rank, const float* slot_w, const void* const* recv,
                                 const void* const* tags, int32_t rank, int32_t P, void* stream_) {
  CSMOE_CHECK_ARG(src && sel && slot_to_row && pad_offsets && dest_base, "csmoe_ep_dispatch: NULL pointer");
  CSMOE_CHECK_ARG(D > 0 && D % 8 == 0 && K >= 1 && experts_per_rank >= 1, "csmoe_ep_dispatch: D %% 8 == 0, K >= 1");
  CSMOE_CHECK_ARG(P >= 1 && P <= kMaxRanks && rank >= 0 && rank < P, "csmoe_ep_dispatch: bad rank / P");
  if (n_slots == 0) return CSMOE_OK;
  Peers rv, tg;
  int rc = fill_peers(rv, recv, P, false, "csmoe_ep_dispatch(recv)");
  if (rc != CSMOE_OK) return rc;
  rc = fill_peers(tg, tags, P, true, "csmoe_ep_dispatch(tags)");
  if (rc != CSMOE_OK) return rc;
  cudaStream_t stream = as_stream(stream_);
  EP_DISPATCH_DTYPE(dtype, (ep_dispatch_kernel<T><<<warp_grid(n_slots), 256, 0, stream>>>(
                               static_cast<const T*>(src), D, K, n_slots, sel, slot_to_row, pad_offsets, dest_base,
                               experts_per_rank, slot_w, rv, tg, rank)));
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_ep_row_ptrs(const int64_t* tags, const int32_t* tile_expert, const int32_t* recv_counts,
                                 const int32_t* recv_pad_offsets, int32_t experts_per_rank, int64_t row_cap,
                                 const void* const* ret, int64_t ret_ld, int32_t dtype, int32_t P, uint64_t* c_rows,
                                 void* recv, int32_t D, void* stream_) {
  CSMOE_CHECK_ARG(tile_expert && recv_counts && recv_pad_offsets, "csmoe_ep_row_ptrs: NULL pointer");
  CSMOE_CHECK_ARG(P >= 1 && P <= kMaxRanks && experts_per_rank >= 1 && row_cap > 0, "csmoe_ep_row_ptrs: bad sizes");
  CSMOE_CHECK_ARG(c_rows == nullptr || (tags != nullptr && ret != nullptr), "csmoe_ep_row_ptrs: c_rows needs tags and ret");
  CSMOE_CHECK_ARG(recv == nullptr || (D > 0 && D % 8 == 0), "csmoe_ep_row_ptrs: D %% 8 == 0");
  Peers rt;
  int rc = fill_peers(rt, ret, P, true, "csmoe_ep_row_ptrs(ret)");
  if (rc != CSMOE_OK) return rc;
  cudaStream_t stream = as_stream(stream_);
  EP_DISPATCH_DTYPE(dtype, (ep_row_ptrs_kernel<T><<<warp_grid(row_cap), 256, 0, stream>>>(
                               reinterpret_cast<const long long*>(tags), tile_expert, recv_counts, recv_pad_offsets,
                               experts_per_rank, row_cap, rt, ret_ld, reinterpret_cast<unsigned long long*>(c_rows),
                               static_cast<T*>(recv), D)));
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_ep_push_rows(const void* src, int32_t dtype, int64_t ld, int32_t D, int64_t rows,
                                  const uint64_t* dst_rows, void* stream_) {
  CSMOE_CHECK_ARG(src && dst_rows, "csmoe_ep_push_rows: NULL pointer");
  CSMOE_CHECK_ARG(D > 0 && D % 8 == 0 && ld >= D, "csmoe_ep_push_rows: D %% 8 == 0, ld >= D");
  if (rows == 0) return CSMOE_OK;
  cudaStream_t stream = as_stream(stream_);
  EP_DISPATCH_DTYPE(dtype, (ep_push_rows_kernel<T><<<warp_grid(rows), 256, 0, stream>>>(
                               static_cast<const T*>(src), ld, D, rows,
                               reinterpret_cast<const unsigned long long*>(dst_rows))));
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_ep_gather_push(const void* src, int32_t src_dtype, int64_t n, const void* const* dst, int32_t dst_dtype,
                                    int64_t dst_offset, int32_t rank, int32_t P, void* stream_) {
  CSMOE_CHECK_ARG(src != nullptr && n >= 0 && n % 8 == 0 && dst_offset >= 0 && dst_offset % 8 == 0,
                  "csmoe_ep_gather_push: n and dst_offset must be multiples of 8");
  CSMOE_CHECK_ARG(P >= 1 && P <= kMaxRanks && rank >= 0 && rank < P, "csmoe_ep_gather_push: bad rank / P");
  if (n == 0) return CSMOE_OK;
  Peers d;
  int rc = fill_peers(d, dst, P, false, "csmoe_ep_gather_push");
  if (rc != CSMOE_OK) return rc;
  cudaStream_t stream = as_stream(stream_);
  const long long vecs = n / 8;
  const long long cap = static_cast<long long>(num_sms() > 0 ? num_sms() : 148) * 8;
  const unsigned grid = static_cast<unsigned>(std::min<long long>((vecs + 255) / 256, cap));
  if (src_dtype == CSMOE_F32 && dst_dtype == CSMOE_BF16) {
    ep_gather_push_kernel<float, __nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const float*>(src), n, d, dst_offset, rank, P);
  } else if (src_dtype == CSMOE_BF16 && dst_dtype == CSMOE_BF16) {
    ep_gather_push_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), n, d, dst_offset, rank, P);
  } else if (src_dtype == CSMOE_F32 && dst_dtype == CSMOE_F32) {
    ep_gather_push_kernel<float, float><<<grid, 256, 0, stream>>>(static_cast<const float*>(src), n, d, dst_offset, rank, P);
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_ep_gather_push: unsupported dtype pair %d -> %d", src_dtype, dst_dtype);
  }
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_ep_reduce_pull(const void* const* src, int64_t src_offset, int64_t n, void* out, int32_t out_dtype,
                                    int32_t P, void* stream_) {
  CSMOE_CHECK_ARG(out != nullptr && n >= 0 && n % 4 == 0 && src_offset >= 0 && src_offset % 4 == 0,
                  "csmoe_ep_reduce_pull: n and src_offset must be multiples of 4");
  CSMOE_CHECK_ARG(P >= 1 && P <= kMaxRanks, "csmoe_ep_reduce_pull: bad P");
  if (n == 0) return CSMOE_OK;
  Peers s;
  int rc = fill_peers(s, src, P, false, "csmoe_ep_reduce_pull");
  if (rc != CSMOE_OK) return rc;
  cudaStream_t stream = as_stream(stream_);
  if (out_dtype == CSMOE_F32) {
    launch_reduce_pull<float>(s, src_offset, n, static_cast<float*>(out), P, stream);
  } else if (out_dtype == CSMOE_BF16) {
    launch_reduce_pull<__nv_bfloat16>(s, src_offset, n, static_cast<__nv_bfloat16*>(out), P, stream);
  } else {
    CSMOE_CHECK_ARG(false, "csmoe_ep_reduce_pull: unsupported output dtype %d", out_dtype);
  }
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}
