// TEST INFRASTRUCTURE ONLY: the SIMT emulator's launcher and the handful of CUDA runtime entry points the C-ABI host
// functions of the emulated kernels call (see simt.h).
#include "simt.h"

#include <ctime>

namespace simt {

dim3 g_grid, g_block;
Barrier g_block_bar;
std::vector<uint8_t> g_dyn_smem;

Ctx& ctx() {
  static thread_local Ctx c;
  return c;
}

// One OS thread per CUDA thread of a block, created once per launch; the blocks run one after another (statics stand in
// for shared memory), separated by a barrier all threads take part in.
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int nthreads = static_cast<int>(block.x * block.y * block.z);
  if (nthreads <= 0 || grid.x * grid.y * grid.z == 0) return;
  g_grid = grid;
  g_block = block;
  g_dyn_smem.assign(smem + 64, 0xCD);      // garbage, like real shared memory
  const int nwarps = (nthreads + 31) / 32;
  std::vector<Warp> warps(nwarps);
  Barrier between;                         // full-strength barrier between blocks (nobody has dropped out of it)
  between.reset(nthreads);
  std::vector<std::thread> pool;
  pool.reserve(nthreads);
  for (int t = 0; t < nthreads; ++t) {
    pool.emplace_back([&, t]() {
      Ctx& c = ctx();
      c.tid = make_uint3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
      c.lane = t % 32;
      c.warp = &warps[t / 32];
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            if (t == 0) {
              g_block_bar.reset(nthreads);
              for (int w = 0; w < nwarps; ++w) {
                const int n = std::min(32, nthreads - 32 * w);
                warps[w].bar.reset(n);
                warps[w].alive = n == 32 ? 0xffffffffu : ((1u << n) - 1u);
              }
            }
            between.wait();
            c.bid = make_uint3(bx, by, bz);
            body();
            c.warp->bar.drop();            // this thread has left the kernel: later barriers do not wait for it
            g_block_bar.drop();
            between.wait();
          }
    });
  }
  for (auto& th : pool) th.join();
}

unsigned long long wall_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return static_cast<unsigned long long>(ts.tv_sec) * 1000000000ull + static_cast<unsigned long long>(ts.tv_nsec);
}

}  // namespace simt

static int sms() {
  const char* v = getenv("SIMT_SMS");
  return v ? atoi(v) : 4;
}

extern "C" {
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "simt: no error"; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
  *v = a == cudaDevAttrMultiProcessorCount ? sms() : (a == cudaDevAttrComputeCapabilityMajor ? 10 : 0);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
// peer memory: the emulated "ranks" share anonymous mappings created by the test; CUDA IPC itself is not emulated
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned int) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaErrorNotSupported; }
}
