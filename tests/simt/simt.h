// TEST INFRASTRUCTURE ONLY -- a small SIMT emulator for the host.
//
// tests/simt_host.py compiles the shipped .cu sources of the bandwidth-bound kernels (routing maps, router, permute /
// combine, losses, competition tail, activations, block tail) with g++ against this header: every CUDA thread of a block
// is an OS thread, blocks run one after another, `__shared__` variables are statics, `__syncthreads()` and the warp
// collectives (`__shfl_*_sync`, `__match_any_sync`, `__syncwarp`) are barriers plus an exchange through per-warp slots.
// The C-ABI entry points are then called with host pointers, so `-m "not gpu"` tests run the kernels' own source -- index
// algebra, reductions orders, rounding points -- against the oracle.  What this cannot see: memory-model effects between
// non-synchronised threads, performance, TMA / tcgen05 (those kernels are not compiled here).  Nothing in the product
// includes this file.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __shared__
#undef __constant__
#undef __forceinline__
#undef __launch_bounds__
#undef __noinline__
#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __constant__ static
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)

namespace simt {

// Barrier whose participant count shrinks when a thread leaves the kernel early (`if (t >= T) return;`).  Sleeping is a
// futex wait on the generation counter (simt_rt.cpp, C++20 atomic wait): no mutex is re-acquired on wake-up, which is
// what made a condition-variable barrier of 256 threads cost milliseconds.
class Barrier {
 public:
  void reset(int n);
  void wait();
  void drop();

 private:
  std::mutex m_;
  int expected_ = 0, waiting_ = 0;
  std::atomic<unsigned> gen_{0};
};

struct Warp {
  Barrier bar;
  uint64_t slot[32];
  unsigned alive;   // lanes that exist in this warp (a partial last warp has fewer than 32)
};

struct Ctx {
  uint3 tid, bid;
  int lane;
  Warp* warp;
};

extern dim3 g_grid, g_block;
extern cudaError_t g_last_error;
extern Barrier g_block_bar;
extern std::vector<uint8_t> g_dyn_smem;
Ctx& ctx();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

inline void* dyn_smem() { return g_dyn_smem.data(); }
unsigned long long wall_ns();

// every participating lane deposits 8 bytes and gets all 32 back
inline void exchange(uint64_t mine, uint64_t (&all)[32]) {
  Ctx& c = ctx();
  c.warp->slot[c.lane] = mine;
  c.warp->bar.wait();
  for (int i = 0; i < 32; ++i) all[i] = c.warp->slot[i];
  c.warp->bar.wait();
}
template <typename T>
inline uint64_t bits(T v) {
  static_assert(sizeof(T) <= 8, "warp collectives move at most 8 bytes");
  uint64_t u = 0;
  memcpy(&u, &v, sizeof(T));
  return u;
}
template <typename T>
inline T from_bits(uint64_t u) {
  T v;
  memcpy(&v, &u, sizeof(T));
  return v;
}

}  // namespace simt

#define threadIdx (simt::ctx().tid)
#define blockIdx (simt::ctx().bid)
#define blockDim (simt::g_block)
#define gridDim (simt::g_grid)

inline void __syncthreads() { simt::g_block_bar.wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { simt::ctx().warp->bar.wait(); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() {}
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

template <typename T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  uint64_t all[32];
  simt::exchange(simt::bits(v), all);
  const int lane = simt::ctx().lane;
  const int base = lane / width * width;
  return simt::from_bits<T>(all[base + (src % width + width) % width]);
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int mask, int width = 32) {
  uint64_t all[32];
  simt::exchange(simt::bits(v), all);
  const int lane = simt::ctx().lane, src = lane ^ mask;
  return (src / width == lane / width) ? simt::from_bits<T>(all[src]) : v;
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
  uint64_t all[32];
  simt::exchange(simt::bits(v), all);
  const int lane = simt::ctx().lane, src = lane - static_cast<int>(delta);
  return (src >= lane / width * width) ? simt::from_bits<T>(all[src]) : v;
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
  uint64_t all[32];
  simt::exchange(simt::bits(v), all);
  const int lane = simt::ctx().lane, src = lane + static_cast<int>(delta);
  return (src < (lane / width + 1) * width) ? simt::from_bits<T>(all[src]) : v;
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  uint64_t all[32];
  simt::exchange(pred ? 1u : 0u, all);
  unsigned r = 0;
  const unsigned alive = simt::ctx().warp->alive & mask;
  for (int i = 0; i < 32; ++i)
    if (((alive >> i) & 1u) && all[i]) r |= 1u << i;
  return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == (simt::ctx().warp->alive & mask); }
template <typename T>
inline unsigned __match_any_sync(unsigned mask, T v) {
  uint64_t all[32];
  simt::exchange(simt::bits(v), all);
  const unsigned alive = simt::ctx().warp->alive & mask;
  const uint64_t mine = simt::bits(v);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i)
    if (((alive >> i) & 1u) && all[i] == mine) r |= 1u << i;
  return r;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz(static_cast<unsigned>(v)); }
template <typename T>
inline T __ldg(const T* p) { return *p; }
inline unsigned __umulhi(unsigned a, unsigned b) { return static_cast<unsigned>((static_cast<uint64_t>(a) * b) >> 32); }
inline float __uint_as_float(unsigned u) { return simt::from_bits<float>(u); }
inline unsigned __float_as_uint(float f) { return static_cast<unsigned>(simt::bits(f)); }
inline float __int_as_float(int u) { return simt::from_bits<float>(static_cast<unsigned>(u)); }
inline int __float_as_int(float f) { return static_cast<int>(simt::bits(f)); }
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F*, cudaFuncAttribute, int) { return cudaSuccess; }   // the C++ overload nvcc provides
inline void __trap() { fprintf(stderr, "simt: __trap()\n"); abort(); }
// glibc declares functions named __expf / __logf: macros, after <cmath>
#define __expf(x) expf(x)
#define __logf(x) logf(x)
#define __fdividef(a, b) ((a) / (b))
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
