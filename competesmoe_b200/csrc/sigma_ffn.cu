// Fused sigma-MoE expert FFN (expert size H = 128): the pretrain plugin's two `cvmm` calls + activation in ONE kernel
// per direction, so that the [rows, H] hidden activations never round-trip through HBM and the K-fold expanded copy of
// the tokens is never written.   Reference: moe_pretrain_model/layers/moe/competesmoe.py:510-522 (compute_moe_main),
// layers/moe/moe.py:397-416 (compute_scores), layers/cvmm.py:61-168 (cvmm_kernel), :194-345 (cvmm_backward_kernel3).
//
//   forward   per 128-row tile of the padded expert-major row space (one expert per tile):
//               S = gather(x)[128, D] . keys[e][D, H]          A rows arrive by TMA gather4 straight from the token-major x
//               h = relu(S + bias[e])  -> bf16 -> shared memory (the A operand of the second MMA) and, once, to HBM
//               y = h[128, H] . values[e][H, Dout]             128-column chunks, double-buffered in TMEM
//             y [row_cap, Dout] leaves through TMA stores; the gate-weighted combine stays a separate deterministic pass.
//   backward  per tile:  dh = gather(dout)[128, Dout] . values[e]^T ;  dw[slot] = <h, dh> ;  dz = w * dh * [h > 0]
//             dxr = dz[128, H] . keys[e]^T  -> [row_cap, D];   also writes dz and hw = w * h for the weight gradients
//   wgrad     C[e] = A[rows of e]^T . gather(B)[rows of e]   (REDUCE over an expert's rows; the gathered operand is read
//             from the token-major tensor through the same gather4 loads): dvalues = hw^T . gather(dout),
//             dkeys^T = dz^T . gather(x)
//
// Warp roles (416 threads, 1 CTA / SM, persistent over tiles): warps 0-3 = producers -- every thread owns one row of the
// tile and copies its 128-byte k-slices with cp.async (16-byte chunks, written in the 128-byte-swizzle pattern the UMMA
// descriptors expect; completion is signalled with cp.async.mbarrier.arrive.noinc on the stage's "full" barrier), thread 0
// also issues the TMA loads of the weight operand; warp 4 = MMA issuer (one lane); warps 5-12 = epilogue (TMEM lane
// quadrant = warp % 4, column half = (warp - 5) / 4).  TMA gather4 loads were measured first (CSMOE_SIGMA_GATHER4=1 keeps
// them selectable): correct, but ~60-90 ns per 4-row instruction per SM, i.e. 1 TB/s over the chip -- 43 us of the 118 us
// forward at the C4 shape -- so the row gather runs on the LSU path instead.  The issuer runs GEMM-1 of tile i+1 before
// GEMM-2 of tile i, so the S -> relu -> shared-memory hop of tile i overlaps tensor work.  Every mbarrier spin is bounded.
#include "common.h"
#include "ptx.cuh"

namespace csmoe {
namespace {

constexpr int kBM = 128;                 // rows per tile
constexpr int kBK = 64;                  // k elements per operand block (one 128-byte swizzle row)
constexpr int kH = 128;                  // expert hidden size this path is specialised for
constexpr int kStages = 4;
constexpr int kStageBytes = 32768;       // GEMM-1: A 128x64 (16 KiB) + B 64x128 (16 KiB);  GEMM-2: B 128x128 (32 KiB)
constexpr int kA2Bytes = kBM * kH * 2;   // relu(S) / dz tile, K-major, two 64-column blocks of 16 KiB
constexpr int kEpiWarps = 8;
constexpr int kProdWarps = 4;            // 128 producer threads: one tile row (or (k-row, sub-tile) pair) each
constexpr int kProdThreads = 32 * kProdWarps;
constexpr int kMmaWarp = kProdWarps;
constexpr int kEpiWarp0 = kProdWarps + 1;
constexpr int kThreads = 32 * (kProdWarps + 1 + kEpiWarps);
constexpr int kStgBytes = 4096;          // per epilogue warp: two 32x32 bf16 staging tiles for the TMA stores
constexpr int kSub = 8192;               // one 64 x 64 bf16 box
constexpr int kSmemBytes = kStages * kStageBytes + 2 * kA2Bytes + kEpiWarps * kStgBytes + 1024 + 256;
constexpr uint32_t kTmemCols = 512;      // S: 2 x 128 columns, second accumulator: 2 x 128 columns

struct Params {
  const int32_t* row_to_slot;   // [row_cap] slot of every padded row, -1 = padding
  const int32_t* tile_expert;   // [row_cap / 128], -1 = no routed rows
  int n_tiles;
  int slots_per_row;            // K: input row of a slot = slot / K
  int D;                        // GEMM-1 contraction (forward: model dim; backward: Dout)
  int Dout;                     // GEMM-2 output columns (forward: Dout; backward: model dim)
  const void* bias;             // forward: [E, H] or NULL
  int bias_fp32;
  int gather;                   // 0 = tiled TMA loads from a pre-gathered copy, 1 = cp.async row gather, 2 = TMA gather4
  const __nv_bfloat16* a_src;   // token-major source of the gathered operand, row pitch = D elements
  // backward only
  const float* slot_w;          // [n_slots] routing weight of every slot (already rounded like the forward pass used it)
  const __nv_bfloat16* h;       // [row_cap, H] saved hidden activations
  float* dw_part;               // [2, n_slots] partial <h, dh> of the two column halves
  long long n_slots;
  int dbg;                      // tuning experiments (CSMOE_SIGMA_DBG bit mask): 1 = no gather copies, 2 = no output-tile stores,
                                // 4 = no saved-tile (h / dz / hw) stores, 16 = output tile by register-direct stores instead of staged TMA stores
  __nv_bfloat16* out;           // [row_cap, Dout] output rows (y / dx rows), row pitch Dout
};

__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int r0, int r1, int r2,
                                            int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// The barrier receives one arrival (counted in its expected total) once every cp.async this thread issued so far has landed.
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 32 x 32 bf16 group (this thread's row = 16 packed words) -> the warp's staging tile (64-byte swizzle) -> one TMA store.
__device__ __forceinline__ void store_group_bf16(uint32_t stg, int lane, const uint32_t (&w)[16], const CUtensorMap* map,
                                                 int col, int row, uint32_t& slot) {
  const uint32_t buf = stg + (slot & 1u) * 2048u;
  ++slot;
  if (lane == 0) ptx::bulk_wait_group_read<1>();
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 4; ++c)
    sts128(buf + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
  ptx::fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(map, buf, col, row);
    ptx::bulk_commit_group();
  }
}

struct Bars {
  uint32_t base;
  __device__ uint32_t full(int s) const { return base + 8u * s; }
  __device__ uint32_t empty(int s) const { return base + 8u * (kStages + s); }
  __device__ uint32_t s_full(int b) const { return base + 8u * (2 * kStages + b); }
  __device__ uint32_t s_empty(int b) const { return base + 8u * (2 * kStages + 2 + b); }
  __device__ uint32_t a2_full(int b) const { return base + 8u * (2 * kStages + 4 + b); }
  __device__ uint32_t a2_empty(int b) const { return base + 8u * (2 * kStages + 6 + b); }
  __device__ uint32_t y_full(int b) const { return base + 8u * (2 * kStages + 8 + b); }
  __device__ uint32_t y_empty(int b) const { return base + 8u * (2 * kStages + 10 + b); }
  __device__ uint32_t tmem_slot() const { return base + 8u * (2 * kStages + 12); }
};

// BWD = false: forward (GEMM-1 B = keys, MN-major; GEMM-2 B = values, MN-major; epilogue 1 = bias + relu)
// BWD = true : backward (GEMM-1 B = values as [H, Dout], K-major; GEMM-2 B = keys as [D, H], K-major; epilogue 1 = dw, dz, hw)
template <bool BWD>
__global__ void __launch_bounds__(kThreads, 1)
sigma_ffn_kernel(const __grid_constant__ CUtensorMap map_a,    // gathered operand: token-major [T, D] (box 64 x 1) or pre-gathered
                 const __grid_constant__ CUtensorMap map_b1,   // fwd keys {H, D, E};  bwd values {Dout, H, E}
                 const __grid_constant__ CUtensorMap map_b2,   // fwd values {Dout, H, E};  bwd keys {H, D, E}
                 const __grid_constant__ CUtensorMap map_t,    // [row_cap, H] tile store: fwd h;  bwd dz
                 const __grid_constant__ CUtensorMap map_t2,   // bwd: hw [row_cap, H]
                 const __grid_constant__ CUtensorMap map_o,    // [row_cap, Dout] 32 x 32 store: fwd y;  bwd dx rows
                 const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a2_base = smem_base + kStages * kStageBytes;
  const uint32_t stg_base = a2_base + 2 * kA2Bytes;
  const Bars bar{stg_base + kEpiWarps * kStgBytes};
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (bar.tmem_slot() - ptx::smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&map_a);
    ptx::prefetch_tmap(&map_b1);
    ptx::prefetch_tmap(&map_b2);
    ptx::prefetch_tmap(&map_o);
  }
  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(bar.full(s), 1 + 32);   // the owning producer warp: lane 0's expect_tx arrive + one arrival per lane
      ptx::mbar_init(bar.empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar.s_full(b), 1);
      ptx::mbar_init(bar.s_empty(b), kEpiWarps);
      ptx::mbar_init(bar.a2_full(b), kEpiWarps);
      ptx::mbar_init(bar.a2_empty(b), 1);
      ptx::mbar_init(bar.y_full(b), 1);
      ptx::mbar_init(bar.y_empty(b), kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == kEpiWarp0) {
    ptx::tmem_alloc(bar.tmem_slot(), kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int nkb1 = p.D / kBK;          // GEMM-1 k blocks
  const int nch = p.Dout / 128;        // GEMM-2 column chunks

  if (warp < kProdWarps) {
    // ===================================================== producers: pipeline iteration `it` (one operand stage) is
    // filled entirely by warp it % 4 -- so a stage's "full" barrier sees 33 arrivals from one warp, and only that warp
    // polls the stage's "empty" barrier.  (All 128 producer threads arriving on / polling one barrier word serialised to
    // profiles/r02k_sigma_bisect.md.)  In pass j lane l copies 16-byte chunk (l & 7) of tile row 32 (l >> 3) + j: eight
    // consecutive lanes cover one contiguous 128-byte slice of a token row, and a lane's 32 rows are consecutive, so their
    // slots come in eight 16-byte loads issued back to back (sixteen dependent scalar loads per k-block cost 3.6 k cycles
    // in the first version of the weight-gradient kernel).
    uint32_t it = 0;
    const int lrow = lane >> 3, lchunk = lane & 7;
    const uint32_t dst0 = lrow * 32 * 128;
    auto load_g1 = [&](int t, int e) {
      int tok[32];
      int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
      const int k = p.slots_per_row;
      if (p.gather == 1) {
        const int4* sp = reinterpret_cast<const int4*>(p.row_to_slot + static_cast<long long>(t) * kBM + lrow * 32);
        int4 s4[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s4[j] = __ldg(sp + j);
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // padding rows (slot -1) read token 0 (finite data): masked in the epilogue
          tok[4 * j] = max(s4[j].x, 0) / k;
          tok[4 * j + 1] = max(s4[j].y, 0) / k;
          tok[4 * j + 2] = max(s4[j].z, 0) / k;
          tok[4 * j + 3] = max(s4[j].w, 0) / k;
        }
      } else if (p.gather == 2) {
        const int4 s4 = __ldg(reinterpret_cast<const int4*>(p.row_to_slot + static_cast<long long>(t) * kBM) + lane);
        r0 = s4.x >= 0 ? s4.x / k : 0;
        r1 = s4.y >= 0 ? s4.y / k : 0;
        r2 = s4.z >= 0 ? s4.z / k : 0;
        r3 = s4.w >= 0 ? s4.w / k : 0;
      }
      for (int kb = 0; kb < nkb1; ++kb, ++it) {
        if ((it & (kProdWarps - 1)) != static_cast<uint32_t>(warp)) continue;
        const uint32_t stage = it % kStages;
        if (lane == 0) ptx::mbar_wait(bar.empty(stage), ((it / kStages) & 1u) ^ 1u);
        __syncwarp();
        const uint32_t fb = bar.full(stage);
        const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + 16384;
        if (lane == 0) {
          ptx::mbar_arrive_expect_tx(fb, p.gather == 1 ? 16384 : kStageBytes);
          if (!BWD) {   // keys[e]: [D, H], MN-major B: two 64(n) x 64(k) boxes
            ptx::tma_load_3d(sb, &map_b1, fb, 0, kb * kBK, e);
            ptx::tma_load_3d(sb + kSub, &map_b1, fb, 64, kb * kBK, e);
          } else {      // values[e]: [H, Dout] = [n, k], K-major B: one 64(k) x 128(n) box
            ptx::tma_load_3d(sb, &map_b1, fb, kb * kBK, 0, e);
          }
          if (p.gather == 0) ptx::tma_load_2d(sa, &map_a, fb, kb * kBK, t * kBM);
        }
        if (p.gather == 1) {
          if (!(p.dbg & 1)) {
            const char* base = reinterpret_cast<const char*>(p.a_src) + kb * 128 + lchunk * 16;
            const long long pitch = static_cast<long long>(p.D) * 2;
#pragma unroll
            for (int j = 0; j < 32; ++j)   // row 32 lrow + j: (row & 7) == (j & 7)
              cp_async16(sa + dst0 + j * 128 + ((lchunk ^ (j & 7)) << 4), base + tok[j] * pitch);
          }
          cp_async_arrive_noinc(fb);
        } else {
          if (p.gather == 2) tma_gather4(sa + lane * 512, &map_a, fb, kb * kBK, r0, r1, r2, r3);
          ptx::mbar_arrive(fb);
        }
      }
    };
    auto load_g2 = [&](int e) {
      for (int nc = 0; nc < nch; ++nc, ++it) {
        if ((it & (kProdWarps - 1)) != static_cast<uint32_t>(warp)) continue;
        const uint32_t stage = it % kStages;
        if (lane == 0) ptx::mbar_wait(bar.empty(stage), ((it / kStages) & 1u) ^ 1u);
        __syncwarp();
        const uint32_t fb = bar.full(stage);
        if (lane == 0) {
          const uint32_t st = smem_base + stage * kStageBytes;
          ptx::mbar_arrive_expect_tx(fb, kStageBytes);
          if (!BWD) {   // values[e][k = 0..128, n = nc*128 ..]: MN-major, per 64-k block two 64(n) x 64(k) boxes
#pragma unroll
            for (int kb2 = 0; kb2 < 2; ++kb2)
#pragma unroll
              for (int j = 0; j < 2; ++j)
                ptx::tma_load_3d(st + kb2 * 16384 + j * kSub, &map_b2, fb, nc * 128 + j * 64, kb2 * kBK, e);
          } else {      // keys[e] as [n = D rows, k = H]: K-major, per 64-k block one 64(k) x 128(n) box
#pragma unroll
            for (int kb2 = 0; kb2 < 2; ++kb2) ptx::tma_load_3d(st + kb2 * 16384, &map_b2, fb, kb2 * kBK, nc * 128, e);
          }
        }
        ptx::mbar_arrive(fb);
      }
    };
    int n_done = 0, prev_e = -1;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      const int e = __ldg(p.tile_expert + t);
      if (e < 0) continue;
      load_g1(t, e);
      if (n_done > 0) load_g2(prev_e);
      prev_e = e;
      ++n_done;
    }
    if (n_done > 0) load_g2(prev_e);
  } else if (warp == kMmaWarp) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t kIdesc1 = ptx::make_idesc_bf16(kBM, kH, false, !BWD);
      constexpr uint32_t kIdesc2 = ptx::make_idesc_bf16(kBM, 128, false, !BWD);
      uint32_t stage = 0, phase = 0, ycount = 0;
      auto advance = [&]() {
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      };
      auto bdesc = [&](uint32_t base, int k) {
        return BWD ? ptx::make_smem_desc_sw128(base + k * 32, 16, 1024) : ptx::make_smem_desc_sw128(base + k * 2048, kSub, 1024);
      };
      auto gemm1 = [&](int it) {
        const int sb = it & 1;
        ptx::mbar_wait(bar.s_empty(sb), ((it >> 1) & 1) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + sb * kH;
        for (int kb = 0; kb < nkb1; ++kb) {
          ptx::mbar_wait(bar.full(stage), phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes, sbm = sa + 16384;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            ptx::umma_f16(d, ptx::make_smem_desc_sw128(sa + k * 32, 16, 1024), bdesc(sbm, k), kIdesc1, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit(bar.empty(stage));
          advance();
        }
        ptx::umma_commit(bar.s_full(sb));
      };
      auto gemm2 = [&](int it) {
        const int ab = it & 1;
        ptx::mbar_wait(bar.a2_full(ab), (it >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t a2 = a2_base + ab * kA2Bytes;
        for (int nc = 0; nc < nch; ++nc) {
          const int yb = ycount & 1;
          ptx::mbar_wait(bar.y_empty(yb), ((ycount >> 1) & 1) ^ 1u);
          ptx::mbar_wait(bar.full(stage), phase);
          ptx::tc_fence_after();
          const uint32_t st = smem_base + stage * kStageBytes;
          const uint32_t d = tmem_base + 2 * kH + yb * 128;
#pragma unroll
          for (int kb2 = 0; kb2 < 2; ++kb2)
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              ptx::umma_f16(d, ptx::make_smem_desc_sw128(a2 + kb2 * 16384 + k * 32, 16, 1024), bdesc(st + kb2 * 16384, k),
                            kIdesc2, (kb2 | k) != 0 ? 1u : 0u);
          ptx::umma_commit(bar.empty(stage));
          ptx::umma_commit(bar.y_full(yb));
          advance();
          ++ycount;
        }
        ptx::umma_commit(bar.a2_empty(ab));
      };
      int n_done = 0;
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        if (__ldg(p.tile_expert + t) < 0) continue;
        gemm1(n_done);
        if (n_done > 0) gemm2(n_done - 1);
        ++n_done;
      }
      if (n_done > 0) gemm2(n_done - 1);
    }
  } else {
    // ===================================================== epilogue warps
    const int q = warp & 3, half = (warp - kEpiWarp0) >> 2;
    const int row = q * 32 + lane;
    const uint32_t stg = stg_base + (warp - kEpiWarp0) * kStgBytes;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    uint32_t slot = 0, ycount = 0;

    auto epi1 = [&](int t, int e, int it) {
      const int sb = it & 1, ab = it & 1;
      const uint32_t par = (it >> 1) & 1;
      ptx::mbar_wait(bar.s_full(sb), par);
      ptx::mbar_wait(bar.a2_empty(ab), par ^ 1u);     // GEMM-2 of tile it-2 no longer reads this buffer
      ptx::tc_fence_after();
      if (lane == 0) ptx::bulk_wait_group_read<0>();  // ... and the tile stores issued from it / from the staging tile are done
      __syncwarp();
      const long long grow = static_cast<long long>(t) * kBM + row;
      const int sl = __ldg(p.row_to_slot + grow);
      const bool valid = sl >= 0;
      const uint32_t a2 = a2_base + ab * kA2Bytes + half * 16384;
      const uint32_t taddr = tmem_base + sb * kH + half * 64 + lane_off;
      uint32_t v0[32], v1[32];
      ptx::tmem_ld_32x32b_x32(taddr, v0);
      ptx::tmem_ld_32x32b_x32(taddr + 32, v1);
      float wslot = 0.f, dot = 0.f;
      if (BWD) wslot = valid ? __ldg(p.slot_w + sl) : 0.f;
      ptx::tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(g == 0 ? v0[i] : v1[i]);
        const int c0 = half * 64 + g * 32;
        uint32_t w[16];
        if (!BWD) {
          if (p.bias != nullptr) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float b[8];
              if (p.bias_fp32)
                load8(reinterpret_cast<const float*>(p.bias) + static_cast<long long>(e) * kH + c0 + c * 8, b);
              else
                load8(reinterpret_cast<const __nv_bfloat16*>(p.bias) + static_cast<long long>(e) * kH + c0 + c * 8, b);
#pragma unroll
              for (int i = 0; i < 8; ++i) f[c * 8 + i] = bf16_round(f[c * 8 + i]) + b[i];   // bf16 cvmm result + fp32 bias (moe.py:397-401)
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = valid ? fmaxf(f[i], 0.f) : 0.f;   // relu; padding rows hold exact zeros
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
        } else {
          // f = dh of the UNWEIGHTED upstream row;  dw += <h, dh>;  dz = w * dh * [h > 0];  hw = w * h (for dvalues)
          float hv[32];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float t8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (valid) load8(p.h + grow * kH + c0 + c * 8, t8);
#pragma unroll
            for (int i = 0; i < 8; ++i) hv[c * 8 + i] = t8[i];
          }
          uint32_t hw[16];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            dot += hv[i] * f[i];
            f[i] = hv[i] > 0.f ? wslot * f[i] : 0.f;
            hv[i] = wslot * hv[i];
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            w[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
            hw[i] = pack_bf16(hv[2 * i], hv[2 * i + 1]);
          }
          // hw: 32 rows x 64 columns through the warp's staging tile (128-byte rows, 128-byte swizzle)
#pragma unroll
          for (int c = 0; c < 4; ++c)
            sts128(stg + lane * 128 + (((g * 4 + c) ^ (lane & 7)) << 4), hw[4 * c], hw[4 * c + 1], hw[4 * c + 2], hw[4 * c + 3]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
          sts128(a2 + row * 128 + (((g * 4 + c) ^ (row & 7)) << 4), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
      }
      if (BWD && valid) p.dw_part[static_cast<long long>(half) * p.n_slots + sl] = dot;
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(bar.s_empty(sb));
        ptx::mbar_arrive(bar.a2_full(ab));
      }
      if (lane == 0 && !(p.dbg & 4)) {
        if (BWD) {   // first, so that the wait_group.read<1> of the next staged store covers it
          tma_store_2d(&map_t2, stg, half * 64, t * kBM + q * 32);
          ptx::bulk_commit_group();
        }
        // this warp's 32 rows x 64 columns of h / dz, straight out of the MMA operand buffer (same 128-byte swizzle)
        tma_store_2d(&map_t, a2 + q * 4096, half * 64, t * kBM + q * 32);
        ptx::bulk_commit_group();
      }
    };
    auto epi2 = [&](int t) {
      const bool row_valid = __ldg(p.row_to_slot + static_cast<long long>(t) * kBM + row) >= 0;
      for (int nc = 0; nc < nch; ++nc) {
        const int yb = ycount & 1;
        ptx::mbar_wait(bar.y_full(yb), (ycount >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + 2 * kH + yb * 128 + half * 64 + lane_off;
        uint32_t v0[32], v1[32];
        ptx::tmem_ld_32x32b_x32(taddr, v0);
        ptx::tmem_ld_32x32b_x32(taddr + 32, v1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar.y_empty(yb));     // the accumulator is in registers: hand the buffer back early
        ++ycount;
        if (p.dbg & 2) continue;
        if (!(p.dbg & 16)) {   // 32 x 32 groups through the warp's staging tile and TMA stores
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = pack_bf16(__uint_as_float(v0[2 * i]), __uint_as_float(v0[2 * i + 1]));
          store_group_bf16(stg, lane, w, &map_o, nc * 128 + half * 64, t * kBM + q * 32, slot);
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = pack_bf16(__uint_as_float(v1[2 * i]), __uint_as_float(v1[2 * i + 1]));
          store_group_bf16(stg, lane, w, &map_o, nc * 128 + half * 64 + 32, t * kBM + q * 32, slot);
        } else if (row_valid) {
          // (CSMOE_SIGMA_DBG=16) register-direct: this thread owns 64 consecutive columns (128 bytes) of its row.  Measured
          // slower at the C4 shape (forward 113 vs 79 us): 32 rows x 16 bytes per warp store instruction
          uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<long long>(t) * kBM + row) * p.Dout + nc * 128 + half * 64);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            dst[c] = make_uint4(pack_bf16(__uint_as_float(v0[8 * c]), __uint_as_float(v0[8 * c + 1])),
                                pack_bf16(__uint_as_float(v0[8 * c + 2]), __uint_as_float(v0[8 * c + 3])),
                                pack_bf16(__uint_as_float(v0[8 * c + 4]), __uint_as_float(v0[8 * c + 5])),
                                pack_bf16(__uint_as_float(v0[8 * c + 6]), __uint_as_float(v0[8 * c + 7])));
            dst[4 + c] = make_uint4(pack_bf16(__uint_as_float(v1[8 * c]), __uint_as_float(v1[8 * c + 1])),
                                    pack_bf16(__uint_as_float(v1[8 * c + 2]), __uint_as_float(v1[8 * c + 3])),
                                    pack_bf16(__uint_as_float(v1[8 * c + 4]), __uint_as_float(v1[8 * c + 5])),
                                    pack_bf16(__uint_as_float(v1[8 * c + 6]), __uint_as_float(v1[8 * c + 7])));
          }
        }
      }
    };
    int n_done = 0, prev_t = -1;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      const int e = __ldg(p.tile_expert + t);
      if (e < 0) continue;
      epi1(t, e, n_done);
      if (n_done > 0) epi2(prev_t);
      prev_t = t;
      ++n_done;
    }
    if (n_done > 0) epi2(prev_t);
    if (lane == 0) ptx::bulk_wait_group<0>();   // shared memory must outlive the TMA stores that read it
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarp0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ------------------------------------------------------------------------------------------------ gathered weight gradient
// C[e] [128, N] = A[rows of e, 128]^T . gather(G)[rows of e, N]      (contraction over the expert's padded rows)
//   A: [row_cap, 128] bf16 in the padded expert-major row space (zero on padding rows): hw or dz
//   G: token-major [T, N] bf16; row r reads token row_to_slot[r] / K (padding rows read token 0 against A's zeros)
// One tile = (expert, 128 output columns); both operands MN-major; fp32 accumulation over the whole row range in one
// CTA: deterministic (the reference's cvmm_backward_kernel3 is split-K with fp32 atomics, cvmm.py:194-345).
constexpr int kWStages = 6;
constexpr int kWSmemBytes = kWStages * kStageBytes + 1024 + 256;

struct WParams {
  const int32_t* row_to_slot;
  const int32_t* pad_offsets;
  int num_experts;
  int n_nblocks;        // N / 128
  int slots_per_row;
  int transpose;        // 1: C stored as [e][n][m] (dkeys [D, H]);  0: [e][m][n] (dvalues [H, Dout])
  int gather;           // 1 = cp.async row gather, 2 = TMA gather4
  const __nv_bfloat16* g_src;   // token-major [T, N]
  int N;
  int dbg;              // tuning experiments (CSMOE_SIGMA_DBG bit mask): 1 = no gather copies, 8 = no epilogue stores
  unsigned long long* stats;   // debug (csmoe_sigma_set_stats): per CTA {producer wait empty, mma wait full, mma wait tempty,
                               // epilogue wait tfull, total, k-blocks}
  int c_fp32;
  void* c;
  long long ldc;
  long long c_expert_stride;
};

__global__ void __launch_bounds__(kThreads, 1)
sigma_wgrad_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_g, const WParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kWStages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kWStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kWStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kWStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&map_a);
    ptx::prefetch_tmap(&map_g);
  }
  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kWStages; ++s) {
      ptx::mbar_init(full_bar(s), 1 + 32);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == kEpiWarp0) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const long long total = static_cast<long long>(p.num_experts) * p.n_nblocks;
  const bool st_on = p.stats != nullptr;
  const long long t_begin = clock64();
  unsigned long long w_acc = 0, w_acc2 = 0, w_acc3 = 0, w_acc4 = 0, n_kb = 0;
  auto timed_wait = [&](uint32_t b, uint32_t par, unsigned long long& acc) {
    if (!st_on) {
      ptx::mbar_wait(b, par);
      return;
    }
    const long long t0 = clock64();
    ptx::mbar_wait(b, par);
    acc += static_cast<unsigned long long>(clock64() - t0);
  };

  if (warp < kProdWarps) {
    // producers: pipeline iteration `it` (one 64-row k-block) is filled entirely by warp it % 4 (see sigma_ffn_kernel).
    // In pass j lane l copies 16-byte chunk (l & 7) of k-row 16 (l >> 3) + j of both 64-column sub-tiles.
    uint32_t it = 0;
    const int lrow = lane >> 3, lchunk = lane & 7;
    const uint32_t dst0 = lrow * 16 * 128;
    const int k = p.slots_per_row;
    int4 s4_next[4];
    bool have_next = false;
    int next_r = -1;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
      const int e = static_cast<int>(t / p.n_nblocks), nb = static_cast<int>(t % p.n_nblocks);
      const int r0 = __ldg(p.pad_offsets + e), r1 = __ldg(p.pad_offsets + e + 1);
      have_next = false;
      for (int r = r0; r < r1; r += kBK, ++it) {
        if ((it & (kProdWarps - 1)) != static_cast<uint32_t>(warp)) continue;
        const long long ta = st_on ? clock64() : 0;
        int tok[16];
        if (p.gather == 1) {
          int4 s4[4];
          if (have_next && next_r == r) {
#pragma unroll
            for (int j = 0; j < 4; ++j) s4[j] = s4_next[j];
          } else {
            const int4* sp = reinterpret_cast<const int4*>(p.row_to_slot + r + lrow * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) s4[j] = __ldg(sp + j);
          }
          // this warp's next k-block of the same tile is kProdWarps blocks further on: fetch its slots now, they arrive
          // while this block's copies are issued and the stage wait is served
          next_r = r + kProdWarps * kBK;
          have_next = next_r < r1;
          if (have_next) {
            const int4* sp = reinterpret_cast<const int4*>(p.row_to_slot + next_r + lrow * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) s4_next[j] = __ldg(sp + j);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            tok[4 * j] = max(s4[j].x, 0) / k;
            tok[4 * j + 1] = max(s4[j].y, 0) / k;
            tok[4 * j + 2] = max(s4[j].z, 0) / k;
            tok[4 * j + 3] = max(s4[j].w, 0) / k;
          }
        }
        const uint32_t stage = it % kWStages;
        if (st_on) {
          int chk = 0;
#pragma unroll
          for (int j = 0; j < 16; ++j) chk += tok[j];
          if (chk == -12345) n_kb += 100;            // forces the index loads to have completed here
          w_acc3 += static_cast<unsigned long long>(clock64() - ta);
        }
        if (lane == 0) timed_wait(empty_bar(stage), ((it / kWStages) & 1u) ^ 1u, w_acc);
        __syncwarp();
        const long long tb = st_on ? clock64() : 0;
        ++n_kb;
        const uint32_t fb = full_bar(stage);
        const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + 16384;
        if (lane == 0) {
          ptx::mbar_arrive_expect_tx(fb, p.gather == 1 ? 16384 : kStageBytes);
          ptx::tma_load_2d(sa, &map_a, fb, 0, r);
          ptx::tma_load_2d(sa + kSub, &map_a, fb, 64, r);
        }
        if (p.gather == 1) {
          if (!(p.dbg & 1)) {
            const char* base = reinterpret_cast<const char*>(p.g_src + nb * 128) + lchunk * 16;
            const long long pitch = static_cast<long long>(p.N) * 2;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const char* src = base + tok[j] * pitch;
              const uint32_t dst = sb + dst0 + j * 128 + ((lchunk ^ (j & 7)) << 4);   // k-row 16 lrow + j
              cp_async16(dst, src);                  // sub-tile 0: columns [nb*128, +64)
              cp_async16(dst + kSub, src + 128);     // sub-tile 1: columns [nb*128 + 64, +64)
            }
          }
          cp_async_arrive_noinc(fb);
        } else {
          if (lane < 32) {
            const int4 s4 = __ldg(reinterpret_cast<const int4*>(p.row_to_slot + r + (lane & 15) * 4));
            const int t0 = s4.x >= 0 ? s4.x / k : 0, t1 = s4.y >= 0 ? s4.y / k : 0;
            const int t2 = s4.z >= 0 ? s4.z / k : 0, t3 = s4.w >= 0 ? s4.w / k : 0;
            tma_gather4(sb + (lane >> 4) * kSub + (lane & 15) * 512, &map_g, fb, nb * 128 + (lane >> 4) * 64, t0, t1, t2, t3);
          }
          ptx::mbar_arrive(fb);
        }
        if (st_on) w_acc4 += static_cast<unsigned long long>(clock64() - tb);
      }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      constexpr uint32_t kIdesc = ptx::make_idesc_bf16(128, 128, true, true);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int e = static_cast<int>(t / p.n_nblocks);
        const int nkb = (__ldg(p.pad_offsets + e + 1) - __ldg(p.pad_offsets + e)) / kBK;
        if (nkb == 0) continue;
        timed_wait(tempty_bar(acc), acc_phase ^ 1u, w_acc2);
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + acc * 128;
        for (int kb = 0; kb < nkb; ++kb) {
          timed_wait(full_bar(stage), phase, w_acc);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + 16384;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            ptx::umma_f16(d, ptx::make_smem_desc_sw128(sa + k * 2048, kSub, 1024),
                          ptx::make_smem_desc_sw128(sb + k * 2048, kSub, 1024), kIdesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit(empty_bar(stage));
          if (++stage == kWStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        ptx::umma_commit(tfull_bar(acc));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    const int q = warp & 3, half = (warp - kEpiWarp0) >> 2;
    const int m = q * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
      const int e = static_cast<int>(t / p.n_nblocks), nb = static_cast<int>(t % p.n_nblocks);
      const bool has_acc = __ldg(p.pad_offsets + e + 1) > __ldg(p.pad_offsets + e);
      uint32_t v0[32], v1[32];
      if (has_acc) {
        timed_wait(tfull_bar(acc), acc_phase, w_acc);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + acc * 128 + half * 64 + (static_cast<uint32_t>(q * 32) << 16);
        ptx::tmem_ld_32x32b_x32(taddr, v0);
        ptx::tmem_ld_32x32b_x32(taddr + 32, v1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v0[i] = v1[i] = 0u;
      }
      const int n0 = nb * 128 + half * 64;
      if (p.dbg & 8) continue;
      if (p.transpose) {
        // C[e][n][m]: for a fixed column the 32 lanes write 32 consecutive elements
        if (p.c_fp32) {
          float* c = reinterpret_cast<float*>(p.c) + e * p.c_expert_stride + static_cast<long long>(n0) * p.ldc + m;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            c[static_cast<long long>(i) * p.ldc] = __uint_as_float(v0[i]);
            c[static_cast<long long>(i + 32) * p.ldc] = __uint_as_float(v1[i]);
          }
        } else {
          __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.c) + e * p.c_expert_stride + static_cast<long long>(n0) * p.ldc + m;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            c[static_cast<long long>(i) * p.ldc] = __float2bfloat16_rn(__uint_as_float(v0[i]));
            c[static_cast<long long>(i + 32) * p.ldc] = __float2bfloat16_rn(__uint_as_float(v1[i]));
          }
        }
      } else {
        float f[8];
        if (p.c_fp32) {
          float* c = reinterpret_cast<float*>(p.c) + e * p.c_expert_stride + static_cast<long long>(m) * p.ldc + n0;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(g < 4 ? v0[g * 8 + i] : v1[(g - 4) * 8 + i]);
            store8(c + g * 8, f);
          }
        } else {
          __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.c) + e * p.c_expert_stride + static_cast<long long>(m) * p.ldc + n0;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(g < 4 ? v0[g * 8 + i] : v1[(g - 4) * 8 + i]);
            store8(c + g * 8, f);
          }
        }
      }
    }
  }
  if (st_on && lane == 0) {
    unsigned long long* o = p.stats + static_cast<long long>(blockIdx.x) * 8;
    if (warp == 0) { o[0] = w_acc; o[5] = n_kb; o[6] = w_acc3; o[7] = w_acc4; }
    if (warp == kMmaWarp) { o[1] = w_acc; o[2] = w_acc2; o[4] = static_cast<unsigned long long>(clock64() - t_begin); }
    if (warp == kEpiWarp0) o[3] = w_acc;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarp0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------ host side
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = []() -> EncodeFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeFn>(p);
  }();
  return fn;
}

// bf16 map: dims / box innermost first; strides (bytes) of dims 1.. ; swizzle 128 or 64 bytes
int encode(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
           const cuuint32_t* box, int swizzle_bytes) {
  EncodeFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return CSMOE_ERR_DRIVER;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return CSMOE_ERR_DRIVER;
  }
  return CSMOE_OK;
}

int gather_mode() {   // 1 = cp.async row gather (default), 2 = TMA gather4 (CSMOE_SIGMA_GATHER4=1)
  static int v = []() {
    const char* s = getenv("CSMOE_SIGMA_GATHER4");
    return (s != nullptr && atoi(s) != 0) ? 2 : 1;
  }();
  return v;
}

unsigned long long* g_stats = nullptr;

int dbg_mask() {
  const char* d = getenv("CSMOE_SIGMA_DBG");
  return d != nullptr ? atoi(d) : 0;
}

int gather_box_rows() {
  static int v = []() {
    const char* s = getenv("CSMOE_GATHER4_BOX_ROWS");
    return s != nullptr ? atoi(s) : 1;
  }();
  return v;
}

int map_2d(CUtensorMap* m, const void* base, long long cols, long long rows, long long ld, int box_c, int box_r, int swz) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_r};
  return encode(m, base, 2, dims, str, box, swz);
}
int map_3d(CUtensorMap* m, const void* base, long long cols, long long rows, long long experts, int box_c, int box_r) {
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)experts};
  cuuint64_t str[2] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * rows * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_r, 1};
  return encode(m, base, 3, dims, str, box, 128);
}

#define CSMOE_TRY(expr)            \
  do {                             \
    int _rc = (expr);              \
    if (_rc != CSMOE_OK) return _rc; \
  } while (0)

template <typename K>
int set_smem(K kern, int bytes) {
  CSMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return CSMOE_OK;
}

}  // namespace
}  // namespace csmoe

using namespace csmoe;

/* debug: per-CTA wait-cycle counters of csmoe_sigma_wgrad land in `buf` (8 x uint64 per CTA, >= 148 CTAs); NULL = off */
extern "C" int csmoe_sigma_set_stats(void* buf) {
  g_stats = static_cast<unsigned long long*>(buf);
  return CSMOE_OK;
}

extern "C" int csmoe_sigma_ffn_supported(int64_t D, int32_t H, int64_t Dout) {
  return (H == kH && D > 0 && D % kBK == 0 && Dout > 0 && Dout % 128 == 0) ? 1 : 0;
}

extern "C" int csmoe_sigma_ffn_fwd(const void* x, int64_t T, int32_t D, int32_t Dout, int32_t E, const void* keys,
                                   const void* values, const void* bias, int32_t bias_dtype, const int32_t* row_to_slot,
                                   const int32_t* tile_expert, int64_t row_cap, int32_t slots_per_row, const void* xp,
                                   void* h, void* y, void* stream_) {
  CSMOE_CHECK_ARG(x && keys && values && row_to_slot && tile_expert && h && y, "csmoe_sigma_ffn_fwd: NULL argument");
  CSMOE_CHECK_ARG(csmoe_sigma_ffn_supported(D, kH, Dout), "csmoe_sigma_ffn_fwd: needs D %% 64 == 0 and Dout %% 128 == 0 (H = 128)");
  CSMOE_CHECK_ARG(T > 0 && E >= 1 && slots_per_row >= 1 && row_cap > 0 && row_cap % kBM == 0, "csmoe_sigma_ffn_fwd: bad sizes");
  CUtensorMap ma, mk, mv, mh, my;
  if (xp != nullptr)
    CSMOE_TRY(map_2d(&ma, xp, D, row_cap, D, 64, 128, 128));
  else
    CSMOE_TRY(map_2d(&ma, x, D, T, D, 64, gather_box_rows(), 128));
  CSMOE_TRY(map_3d(&mk, keys, kH, D, E, 64, 64));          // keys [E, D, H]: MN-major B, 64(n) x 64(k) boxes
  CSMOE_TRY(map_3d(&mv, values, Dout, kH, E, 64, 64));     // values [E, H, Dout]: MN-major B
  CSMOE_TRY(map_2d(&mh, h, kH, row_cap, kH, 64, 32, 128));
  CSMOE_TRY(map_2d(&my, y, Dout, row_cap, Dout, 32, 32, 64));
  Params p{};
  p.row_to_slot = row_to_slot;
  p.tile_expert = tile_expert;
  p.n_tiles = static_cast<int>(row_cap / kBM);
  p.slots_per_row = slots_per_row;
  p.D = D;
  p.Dout = Dout;
  p.bias = bias;
  p.bias_fp32 = bias_dtype == CSMOE_F32;
  p.gather = xp == nullptr ? gather_mode() : 0;
  p.a_src = static_cast<const __nv_bfloat16*>(x);
  p.out = static_cast<__nv_bfloat16*>(y);
  p.dbg = dbg_mask();
  static bool configured = false;
  if (!configured) {
    CSMOE_TRY(set_smem(sigma_ffn_kernel<false>, kSmemBytes));
    configured = true;
  }
  const int sms = num_sms() > 0 ? num_sms() : 148;
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
  sigma_ffn_kernel<false><<<grid, kThreads, kSmemBytes, as_stream(stream_)>>>(ma, mk, mv, mh, mh, my, p);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_sigma_ffn_bwd(const void* dout, int64_t T, int32_t D, int32_t Dout, int32_t E, const void* keys,
                                   const void* values, const int32_t* row_to_slot, const int32_t* tile_expert,
                                   int64_t row_cap, int32_t slots_per_row, const float* slot_w, int64_t n_slots,
                                   const void* h, const void* dyp, void* dz, void* hw, void* dxr, float* dw_part,
                                   void* stream_) {
  CSMOE_CHECK_ARG(dout && keys && values && row_to_slot && tile_expert && slot_w && h && dz && hw && dxr && dw_part,
                  "csmoe_sigma_ffn_bwd: NULL argument");
  CSMOE_CHECK_ARG(csmoe_sigma_ffn_supported(D, kH, Dout), "csmoe_sigma_ffn_bwd: needs D %% 64 == 0 and Dout %% 128 == 0 (H = 128)");
  CSMOE_CHECK_ARG(T > 0 && E >= 1 && slots_per_row >= 1 && row_cap > 0 && row_cap % kBM == 0 && D % 128 == 0 && Dout % kBK == 0,
                  "csmoe_sigma_ffn_bwd: bad sizes (needs D %% 128 == 0)");
  CUtensorMap ma, mv, mk, mz, mw, mo;
  if (dyp != nullptr)
    CSMOE_TRY(map_2d(&ma, dyp, Dout, row_cap, Dout, 64, 128, 128));
  else
    CSMOE_TRY(map_2d(&ma, dout, Dout, T, Dout, 64, gather_box_rows(), 128));
  CSMOE_TRY(map_3d(&mv, values, Dout, kH, E, 64, 128));    // values [E, H, Dout] as [n = H, k = Dout]: K-major B
  CSMOE_TRY(map_3d(&mk, keys, kH, D, E, 64, 128));         // keys [E, D, H] as [n = D, k = H]: K-major B
  CSMOE_TRY(map_2d(&mz, dz, kH, row_cap, kH, 64, 32, 128));
  CSMOE_TRY(map_2d(&mw, hw, kH, row_cap, kH, 64, 32, 128));
  CSMOE_TRY(map_2d(&mo, dxr, D, row_cap, D, 32, 32, 64));
  Params p{};
  p.row_to_slot = row_to_slot;
  p.tile_expert = tile_expert;
  p.n_tiles = static_cast<int>(row_cap / kBM);
  p.slots_per_row = slots_per_row;
  p.D = Dout;        // GEMM-1 contracts over the layer's output dimension
  p.Dout = D;        // GEMM-2 produces the d x rows
  p.gather = dyp == nullptr ? gather_mode() : 0;
  p.a_src = static_cast<const __nv_bfloat16*>(dout);
  p.out = static_cast<__nv_bfloat16*>(dxr);
  p.dbg = dbg_mask();
  p.slot_w = slot_w;
  p.h = static_cast<const __nv_bfloat16*>(h);
  p.dw_part = dw_part;
  p.n_slots = n_slots;
  static bool configured = false;
  if (!configured) {
    CSMOE_TRY(set_smem(sigma_ffn_kernel<true>, kSmemBytes));
    configured = true;
  }
  const int sms = num_sms() > 0 ? num_sms() : 148;
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
  sigma_ffn_kernel<true><<<grid, kThreads, kSmemBytes, as_stream(stream_)>>>(ma, mv, mk, mz, mw, mo, p);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

extern "C" int csmoe_sigma_wgrad(const void* a, const void* g, int64_t T, int32_t N, int32_t E, const int32_t* row_to_slot,
                                 const int32_t* pad_offsets, int64_t row_cap, int32_t slots_per_row, int32_t transpose,
                                 void* c, int32_t c_dtype, void* stream_) {
  CSMOE_CHECK_ARG(a && g && row_to_slot && pad_offsets && c, "csmoe_sigma_wgrad: NULL argument");
  CSMOE_CHECK_ARG(T > 0 && N > 0 && N % 128 == 0 && E >= 1 && row_cap > 0 && row_cap % kBM == 0 && slots_per_row >= 1,
                  "csmoe_sigma_wgrad: bad sizes (N must be a multiple of 128)");
  CSMOE_CHECK_ARG(c_dtype == CSMOE_F32 || c_dtype == CSMOE_BF16, "csmoe_sigma_wgrad: bad output dtype");
  CUtensorMap ma, mg;
  CSMOE_TRY(map_2d(&ma, a, kH, row_cap, kH, 64, 64, 128));
  CSMOE_TRY(map_2d(&mg, g, N, T, N, 64, gather_box_rows(), 128));
  WParams p{};
  p.row_to_slot = row_to_slot;
  p.pad_offsets = pad_offsets;
  p.num_experts = E;
  p.n_nblocks = N / 128;
  p.slots_per_row = slots_per_row;
  p.transpose = transpose;
  p.gather = gather_mode();
  p.g_src = static_cast<const __nv_bfloat16*>(g);
  p.N = N;
  p.dbg = dbg_mask();
  p.stats = g_stats;
  p.c_fp32 = c_dtype == CSMOE_F32;
  p.c = c;
  p.ldc = transpose ? kH : N;
  p.c_expert_stride = static_cast<long long>(kH) * N;
  static bool configured = false;
  if (!configured) {
    CSMOE_TRY(set_smem(sigma_wgrad_kernel, kWSmemBytes));
    configured = true;
  }
  const int sms = num_sms() > 0 ? num_sms() : 148;
  const long long total = static_cast<long long>(E) * p.n_nblocks;
  const int grid = static_cast<int>(total < sms ? total : sms);
  sigma_wgrad_kernel<<<grid, kThreads, kWSmemBytes, as_stream(stream_)>>>(ma, mg, p);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}
