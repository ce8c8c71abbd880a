// Routing metadata: stable counting sort of the flattened expert ids (histogram -> scan -> stable scatter).
// Integer work, bit-exact by construction.  Replaces `fsel.sort()` + index arithmetic of cvmm_prepare_sel2
// (moe_pretrain_model/layers/cvmm.py:580-592) and the E x torch.where of compute_moe (moe_model/model/moe/moe.py:189-191).
//
// Layout produced: the "padded expert-major row space": expert e owns rows [pad_offsets[e], pad_offsets[e] + counts[e]),
// pad_offsets are multiples of CSMOE_ROW_TILE so a GEMM row tile never straddles experts.
#include "common.h"

namespace csmoe {
namespace {

constexpr int kChunk = 2048;      // slots per CTA
constexpr int kThreads = 256;     // 8 warps, each owns kChunk/8 = 256 consecutive slots
constexpr int kWarps = kThreads / 32;
constexpr int kPerWarp = kChunk / kWarps;
constexpr int kMaxExperts = 1024;

// Expert ids come from this library's own top-k kernels or, through the public cvmm_prepare_sel, from the caller.  An id
// outside [0, E) is clamped into range so that a bad index can never become an out-of-bounds shared / global write here
// or in the gather / combine kernels that consume the maps (the reference would raise a device-side assert instead).
__device__ __forceinline__ int clamp_expert(int e, int E) { return e < 0 ? 0 : (e >= E ? E - 1 : e); }

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Phase A: per-chunk, per-expert histogram (no atomics: each warp walks its sub-chunk in order, match_any groups lanes).
// Also clears row_to_slot / tile_expert with a grid-stride fill.
__global__ void __launch_bounds__(kThreads)
route_hist_kernel(const int32_t* __restrict__ sel, long long n_slots, int E, int32_t* __restrict__ chunk_counts,
                  int32_t* __restrict__ row_to_slot, long long row_cap, int32_t* __restrict__ tile_expert) {
  extern __shared__ int32_t sh[];  // [kWarps][E]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kWarps * E; i += kThreads) sh[i] = 0;
  __syncthreads();
  const long long base = static_cast<long long>(blockIdx.x) * kChunk + static_cast<long long>(warp) * kPerWarp;
  int32_t* mine = sh + warp * E;
  for (int it = 0; it < kPerWarp / 32; ++it) {
    const long long j = base + it * 32 + lane;
    const bool ok = j < n_slots;
    const int e = ok ? clamp_expert(sel[j], E) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, e);
    if (ok && (peers & lanemask_lt()) == 0) mine[e] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += kThreads) {
    int s = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += sh[w * E + e];
    chunk_counts[static_cast<long long>(blockIdx.x) * E + e] = s;
  }
  if (row_to_slot != nullptr) {
    for (long long r = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; r < row_cap;
         r += static_cast<long long>(gridDim.x) * kThreads)
      row_to_slot[r] = -1;
  }
  if (tile_expert != nullptr) {
    const long long n_tiles = row_cap / CSMOE_ROW_TILE;
    for (long long r = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; r < n_tiles;
         r += static_cast<long long>(gridDim.x) * kThreads)
      tile_expert[r] = -1;
  }
}

// Phase B (one CTA): per-expert totals, exclusive scans over experts (plain and padded) and, per expert, the running
// base of every chunk.  chunk_counts[c][e] is overwritten with the exclusive prefix over chunks.
__global__ void __launch_bounds__(1024)
route_scan_kernel(int32_t* __restrict__ chunk_counts, int n_chunks, int E, int row_tile, int32_t* __restrict__ counts,
                  int32_t* __restrict__ offsets, int32_t* __restrict__ pad_offsets, int32_t* __restrict__ tile_expert) {
  __shared__ int32_t s_cnt[kMaxExperts];
  __shared__ int32_t s_off[kMaxExperts + 1];
  __shared__ int32_t s_pad[kMaxExperts + 1];
  // one warp per expert: the chunks are scanned 32 at a time with a warp prefix sum (a thread per expert walking the
  // chunks one by one was a chain of dependent global loads: 16.7 us at 64 experts x 32 chunks)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  for (int e = warp; e < E; e += n_warps) {
    int run = 0;
    for (int c0 = 0; c0 < n_chunks; c0 += 32) {
      const int c = c0 + lane;
      const long long i = static_cast<long long>(c) * E + e;
      const int v = c < n_chunks ? chunk_counts[i] : 0;
      int x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (c < n_chunks) chunk_counts[i] = run + x - v;
      run += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) {
      s_cnt[e] = run;
      counts[e] = run;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int o = 0, po = 0;
    for (int e = 0; e < E; ++e) {
      s_off[e] = o;
      s_pad[e] = po;
      o += s_cnt[e];
      po += (s_cnt[e] + row_tile - 1) / row_tile * row_tile;
    }
    s_off[E] = o;
    s_pad[E] = po;
  }
  __syncthreads();
  for (int e = threadIdx.x; e <= E; e += blockDim.x) {
    offsets[e] = s_off[e];
    pad_offsets[e] = s_pad[e];
  }
  if (tile_expert != nullptr) {
    for (int e = 0; e < E; ++e) {
      const int t0 = s_pad[e] / CSMOE_ROW_TILE, t1 = s_pad[e + 1] / CSMOE_ROW_TILE;
      for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) tile_expert[t] = e;
    }
  }
}

// Phase C: stable scatter.  Position of slot j = offsets[e] + (#slots with the same expert before j).
__global__ void __launch_bounds__(kThreads)
route_scatter_kernel(const int32_t* __restrict__ sel, long long n_slots, int E, const int32_t* __restrict__ chunk_base,
                     const int32_t* __restrict__ offsets, const int32_t* __restrict__ pad_offsets,
                     int32_t* __restrict__ sorted_sel, int64_t* __restrict__ sort_index, int32_t* __restrict__ slot_to_row,
                     int32_t* __restrict__ row_to_slot) {
  extern __shared__ int32_t sh[];  // [kWarps][E] running positions (relative to the expert segment start)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kWarps * E; i += kThreads) sh[i] = 0;
  __syncthreads();
  const long long base = static_cast<long long>(blockIdx.x) * kChunk + static_cast<long long>(warp) * kPerWarp;
  int32_t* mine = sh + warp * E;
  // pass 1: per-warp histogram of the sub-chunk
  for (int it = 0; it < kPerWarp / 32; ++it) {
    const long long j = base + it * 32 + lane;
    const bool ok = j < n_slots;
    const int e = ok ? clamp_expert(sel[j], E) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, e);
    if (ok && (peers & lanemask_lt()) == 0) mine[e] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // exclusive prefix over warps, seeded with this chunk's base inside the expert segment
  for (int e = threadIdx.x; e < E; e += kThreads) {
    int run = chunk_base[static_cast<long long>(blockIdx.x) * E + e];
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const int v = sh[w * E + e];
      sh[w * E + e] = run;
      run += v;
    }
  }
  __syncthreads();
  // pass 2: assign positions in slot order
  for (int it = 0; it < kPerWarp / 32; ++it) {
    const long long j = base + it * 32 + lane;
    const bool ok = j < n_slots;
    const int e = ok ? clamp_expert(sel[j], E) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, e);
    if (ok) {
      const int rank = mine[e] + __popc(peers & lanemask_lt());
      const int pos = offsets[e] + rank;
      const int row = pad_offsets[e] + rank;
      if (sorted_sel) sorted_sel[pos] = e;
      if (sort_index) sort_index[pos] = j;
      if (slot_to_row) slot_to_row[j] = row;
      if (row_to_slot) row_to_slot[row] = static_cast<int32_t>(j);
    }
    __syncwarp();
    if (ok && (peers & lanemask_lt()) == 0) mine[e] += __popc(peers);
    __syncwarp();
  }
}

}  // namespace
}  // namespace csmoe

using namespace csmoe;

extern "C" int64_t csmoe_route_row_cap(int64_t n_slots, int32_t num_experts, int32_t row_tile) {
  if (n_slots < 0 || num_experts <= 0 || (row_tile != 128 && row_tile != 256)) return -1;
  const int64_t worst = n_slots + static_cast<int64_t>(num_experts) * (row_tile - 1);
  return (worst + row_tile - 1) / row_tile * row_tile;
}

extern "C" int64_t csmoe_route_workspace_bytes(int64_t n_slots, int32_t num_experts) {
  if (n_slots < 0 || num_experts <= 0) return -1;
  const int64_t n_chunks = (n_slots + kChunk - 1) / kChunk;
  return (n_chunks > 0 ? n_chunks : 1) * num_experts * static_cast<int64_t>(sizeof(int32_t));
}

extern "C" int csmoe_route_build(const int32_t* sel, int64_t n_slots, int32_t num_experts, int32_t row_tile, int64_t row_cap,
                                 int32_t* counts, int32_t* offsets, int32_t* pad_offsets, int32_t* sorted_sel,
                                 int64_t* sort_index, int32_t* slot_to_row, int32_t* row_to_slot, int32_t* tile_expert,
                                 void* workspace, void* stream_) {
  CSMOE_CHECK_ARG(num_experts >= 1 && num_experts <= kMaxExperts, "csmoe_route_build: num_experts must be in [1, %d]",
                  kMaxExperts);
  CSMOE_CHECK_ARG(n_slots >= 0 && n_slots < (1ll << 31), "csmoe_route_build: n_slots out of range");
  CSMOE_CHECK_ARG(counts && offsets && pad_offsets && workspace, "csmoe_route_build: counts/offsets/pad_offsets/workspace are required");
  CSMOE_CHECK_ARG(n_slots == 0 || sel != nullptr, "csmoe_route_build: sel is NULL");
  CSMOE_CHECK_ARG(row_tile == 128 || row_tile == 256, "csmoe_route_build: row_tile must be 128 or 256");
  CSMOE_CHECK_ARG(row_cap % row_tile == 0 && row_cap >= csmoe_route_row_cap(n_slots, num_experts, row_tile),
                  "csmoe_route_build: row_cap %lld too small (need %lld)", (long long)row_cap,
                  (long long)csmoe_route_row_cap(n_slots, num_experts, row_tile));
  cudaStream_t stream = as_stream(stream_);
  const int n_chunks = static_cast<int>((n_slots + kChunk - 1) / kChunk);
  int32_t* chunk_counts = reinterpret_cast<int32_t*>(workspace);
  const size_t smem = static_cast<size_t>(kWarps) * num_experts * sizeof(int32_t);
  const int grid = n_chunks > 0 ? n_chunks : 1;
  route_hist_kernel<<<grid, kThreads, smem, stream>>>(sel, n_slots, num_experts, chunk_counts, row_to_slot, row_cap,
                                                      tile_expert);
  CSMOE_CHECK_LAUNCH();
  route_scan_kernel<<<1, 1024, 0, stream>>>(chunk_counts, grid, num_experts, row_tile, counts, offsets, pad_offsets,
                                            tile_expert);
  CSMOE_CHECK_LAUNCH();
  if (n_chunks > 0) {
    route_scatter_kernel<<<n_chunks, kThreads, smem, stream>>>(sel, n_slots, num_experts, chunk_counts, offsets,
                                                                pad_offsets, sorted_sel, sort_index, slot_to_row,
                                                                row_to_slot);
    CSMOE_CHECK_LAUNCH();
  }
  return CSMOE_OK;
}
