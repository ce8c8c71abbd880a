// TEST INFRASTRUCTURE ONLY: the SIMT emulator's launcher and the handful of CUDA runtime entry points the C-ABI host
// functions of the emulated kernels call (see simt.h).
#include "simt.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <semaphore.h>

#include <ctime>
#include <map>

namespace simt {

dim3 g_grid, g_block;
Barrier g_block_bar;
std::vector<uint8_t> g_dyn_smem;

Ctx& ctx() {
  static thread_local Ctx c;
  return c;
}

void Barrier::reset(int n) {
  std::lock_guard<std::mutex> lk(m_);
  expected_ = n;
  waiting_ = 0;
}
void Barrier::wait() {
  unsigned g;
  bool last = false;
  {
    std::lock_guard<std::mutex> lk(m_);
    g = gen_.load(std::memory_order_relaxed);
    if (++waiting_ >= expected_) {
      waiting_ = 0;
      last = true;
      gen_.store(g + 1, std::memory_order_release);
    }
  }
  if (last) {
    gen_.notify_all();
  } else {
    while (gen_.load(std::memory_order_acquire) == g) gen_.wait(g, std::memory_order_acquire);
  }
}
void Barrier::drop() {
  bool release = false;
  {
    std::lock_guard<std::mutex> lk(m_);
    --expected_;
    if (expected_ > 0 && waiting_ >= expected_) {
      waiting_ = 0;
      release = true;
      gen_.fetch_add(1, std::memory_order_release);
    }
  }
  if (release) gen_.notify_all();
}

// One OS thread per CUDA thread of a block, kept in a pool across launches (creating 256 threads per launch cost 12 ms);
// the blocks of a launch run one after another (statics stand in for shared memory), separated by a barrier all threads
// of the launch take part in.
namespace {
struct Job {
  dim3 grid, block;
  int nthreads = 0, nwarps = 0;
  const std::function<void()>* body = nullptr;
  std::vector<Warp>* warps = nullptr;
  Barrier between;
};
struct Worker {
  std::thread th;
  sem_t go;
};
Job g_job;
std::vector<Worker*> g_pool;          // never destroyed: the threads live until the process exits
pid_t g_pool_pid = 0;
std::atomic<int> g_remaining{0};
sem_t g_done;

void run_blocks(int t) {
  Job& j = g_job;
  Ctx& c = ctx();
  c.tid = make_uint3(t % j.block.x, (t / j.block.x) % j.block.y, t / (j.block.x * j.block.y));
  c.lane = t % 32;
  c.warp = &(*j.warps)[t / 32];
  for (unsigned bz = 0; bz < j.grid.z; ++bz)
    for (unsigned by = 0; by < j.grid.y; ++by)
      for (unsigned bx = 0; bx < j.grid.x; ++bx) {
        if (t == 0) {
          g_block_bar.reset(j.nthreads);
          for (int w = 0; w < j.nwarps; ++w) {
            const int n = std::min(32, j.nthreads - 32 * w);
            (*j.warps)[w].bar.reset(n);
            (*j.warps)[w].alive = n == 32 ? 0xffffffffu : ((1u << n) - 1u);
          }
        }
        j.between.wait();
        c.bid = make_uint3(bx, by, bz);
        (*j.body)();
        c.warp->bar.drop();            // this thread has left the kernel: later barriers do not wait for it
        g_block_bar.drop();
        j.between.wait();
      }
}

void worker_loop(Worker* w, int t) {
  for (;;) {
    while (sem_wait(&w->go) != 0) {
    }
    run_blocks(t);
    if (g_remaining.fetch_sub(1) == 1) sem_post(&g_done);
  }
}
}  // namespace

cudaError_t g_last_error = cudaSuccess;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int nthreads = static_cast<int>(block.x * block.y * block.z);
  // what the CUDA runtime rejects, the emulator rejects: an empty grid or block, more than 1024 threads per block, more
  // dynamic shared memory than an sm_100 block can opt in to, grid.y / grid.z beyond 65535
  if (nthreads <= 0 || nthreads > 1024 || grid.x == 0 || grid.y == 0 || grid.z == 0 || grid.y > 65535 || grid.z > 65535 ||
      smem > 227 * 1024) {
    g_last_error = cudaErrorInvalidConfiguration;
    return;
  }
  if (g_pool_pid != getpid()) {          // first launch, or a forked child: the parent's pool threads do not exist here
    g_pool.clear();
    sem_init(&g_done, 0, 0);
    g_pool_pid = getpid();
  }
  while (static_cast<int>(g_pool.size()) < nthreads) {
    Worker* w = new Worker;
    sem_init(&w->go, 0, 0);
    const int t = static_cast<int>(g_pool.size());
    g_pool.push_back(w);
    w->th = std::thread(worker_loop, w, t);
    w->th.detach();
  }
  g_grid = grid;
  g_block = block;
  g_dyn_smem.assign(smem + 64, 0xCD);      // garbage, like real shared memory
  std::vector<Warp> warps((nthreads + 31) / 32);
  g_job.grid = grid;
  g_job.block = block;
  g_job.nthreads = nthreads;
  g_job.nwarps = static_cast<int>(warps.size());
  g_job.body = &body;
  g_job.warps = &warps;
  g_job.between.reset(nthreads);
  g_remaining.store(nthreads);
  for (int t = 0; t < nthreads; ++t) sem_post(&g_pool[t]->go);
  while (sem_wait(&g_done) != 0) {
  }
}

unsigned long long wall_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return static_cast<unsigned long long>(ts.tv_sec) * 1000000000ull + static_cast<unsigned long long>(ts.tv_nsec);
}

}  // namespace simt

namespace {
struct Seg {
  char name[48];
  size_t bytes;
  bool owner;
};
std::map<void*, Seg> g_segs;
std::mutex g_seg_mu;
int g_seg_counter = 0;
}  // namespace

static int sms() {
  const char* v = getenv("SIMT_SMS");
  return v ? atoi(v) : 4;
}

extern "C" {
cudaError_t cudaGetLastError(void) {
  const cudaError_t e = simt::g_last_error;
  simt::g_last_error = cudaSuccess;
  return e;
}
cudaError_t cudaPeekAtLastError(void) { return simt::g_last_error; }
const char* cudaGetErrorString(cudaError_t e) {
  return e == cudaSuccess ? "no error" : (e == cudaErrorInvalidConfiguration ? "invalid configuration argument (simt)" : "error (simt)");
}
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
// "Device" allocations are POSIX shared-memory segments, so that csmoe_ep_alloc's buffers can be mapped by the other
// emulated ranks (other processes) through the CUDA-IPC calls below: handle = (segment name, size).
cudaError_t cudaMalloc(void** p, size_t n) {
  std::lock_guard<std::mutex> lk(g_seg_mu);
  Seg s;
  snprintf(s.name, sizeof(s.name), "/csmoe_simt_%d_%d", static_cast<int>(getpid()), g_seg_counter++);
  s.bytes = n;
  s.owner = true;
  const int fd = shm_open(s.name, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0 || ftruncate(fd, static_cast<off_t>(n)) != 0) return cudaErrorMemoryAllocation;
  void* q = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (q == MAP_FAILED) return cudaErrorMemoryAllocation;
  g_segs[q] = s;
  *p = q;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
  std::lock_guard<std::mutex> lk(g_seg_mu);
  auto it = g_segs.find(p);
  if (it == g_segs.end()) return cudaErrorInvalidValue;
  munmap(p, it->second.bytes);
  if (it->second.owner) shm_unlink(it->second.name);
  g_segs.erase(it);
  return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
  std::lock_guard<std::mutex> lk(g_seg_mu);
  auto it = g_segs.find(p);
  if (it == g_segs.end()) return cudaErrorInvalidValue;
  memset(h, 0, sizeof(*h));
  memcpy(h->reserved, it->second.name, sizeof(it->second.name));
  memcpy(h->reserved + 48, &it->second.bytes, sizeof(size_t));
  return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned int) {
  std::lock_guard<std::mutex> lk(g_seg_mu);
  Seg s;
  memcpy(s.name, h.reserved, sizeof(s.name));
  s.name[sizeof(s.name) - 1] = 0;
  memcpy(&s.bytes, h.reserved + 48, sizeof(size_t));
  s.owner = false;
  const int fd = shm_open(s.name, O_RDWR, 0);
  if (fd < 0) return cudaErrorInvalidValue;
  void* q = mmap(nullptr, s.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (q == MAP_FAILED) return cudaErrorInvalidValue;
  g_segs[q] = s;
  *p = q;
  return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p) { return cudaFree(p); }
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
}
