// Grouped GEMM for the expert FFNs: persistent, warp-specialised, TMA -> shared memory (128B swizzle) -> tcgen05.mma
// with fp32 accumulators in TMEM -> tcgen05.ld epilogue.  One CTA per SM, 320 threads:
//   warp 0      TMA producer (one lane)
//   warp 1      MMA issuer   (one lane)
//   warps 2..9  epilogue (TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4); warp 2 owns the TMEM allocation
// Two accumulator stages in TMEM (2 x BN columns) let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Modes (see csmoe.h): ROWS  C[rows,n]  = A[rows,k] . B[e]      (A K-major; B K-major [n,k] or MN-major [k,n])
//                      REDUCE C[e][m,n] = A[rows_e,m]^T . B[rows_e,n]   (both operands MN-major)
// The functions replaced are cvmm_kernel / cvmm_backward_kernel3 (moe_pretrain_model/layers/cvmm.py:61-168,194-345) and
// the per-expert nn.Linear calls of moe_model/model/moe/moe.py:196-204.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace csmoe {
namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kEpiWarps = 8;              // two warps per TMEM lane quadrant, each owning half of the tile columns
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kAccStages = 2;
constexpr int kABytes = kBM * kBK * 2;  // 16 KiB
constexpr int kSubTileBytes = 64 * kBK * 2;  // one 64(mn) x 64(k) MN-major box = 8 KiB

struct KParams {
  int n;              // valid output columns
  int m_valid;        // REDUCE: valid output rows per expert
  int num_kb;         // ROWS: K blocks per tile
  int num_experts;
  int num_m_blocks;
  int num_n_blocks;
  int dense;
  int dense_mblocks;  // dense_rows / 128
  int dense_kblocks;  // dense_rows / 64
  int a_expert_rows;
  int b_expert_rows;
  int act;
  int c_fp32;
  int bias_fp32;
  int accumulate;
  int glu_f;          // GLU epilogues: F (forward: n == 2F and tiles interleave gate|up; backward: n == F)
  int epi;            // kEpiPlain / kEpiGluFwd / kEpiActBwd / kEpiGluBwd
  int band;           // n-blocks per rasterisation band
  int num_m_pairs;    // CTA-pair kernel: number of 256-row blocks
  const void* aux;    // backward epilogues: the saved pre-activation z
  long long ldaux;
  const unsigned long long* c_rows;  // plain ROWS epilogue: per-row destination address (expert-parallel return), 0 = skip
  const int* tile_expert;
  const int* pad_offsets;
  void* c;
  void* preact;
  const void* bias;
  long long ldc;
  long long ldpre;
  long long c_expert_stride;
  long long total_tiles;
};

constexpr int kEpiPlain = 0, kEpiGluFwd = 1, kEpiActBwd = 2, kEpiGluBwd = 3;

struct Tile {
  int e, mb, nb, a_row, b_row, nkb;
  bool valid;
};

template <int MODE>
__device__ __forceinline__ Tile decode_tile(const KParams& p, long long t) {
  Tile ti;
  if (MODE == CSMOE_GEMM_ROWS) {
    // Bands of `band` n-blocks, m-blocks swept inside a band, n fastest: the ~148 tiles in flight form a roughly
    // square patch of the output, so both operands are re-used out of L2 while they are hot.
    const long long band_tiles = static_cast<long long>(p.band) * p.num_m_blocks;
    const int b = static_cast<int>(t / band_tiles);
    const int r = static_cast<int>(t % band_tiles);
    const int nb0 = b * p.band;
    const int w = min(p.band, p.num_n_blocks - nb0);
    ti.mb = r / w;
    ti.nb = nb0 + r % w;
    if (p.dense) {
      ti.e = ti.mb / p.dense_mblocks;
      ti.a_row = (ti.mb % p.dense_mblocks) * kBM + ti.e * p.a_expert_rows;
    } else {
      ti.e = __ldg(p.tile_expert + ti.mb);
      ti.a_row = ti.mb * kBM;
    }
    ti.b_row = 0;
    ti.nkb = p.num_kb;
    ti.valid = ti.e >= 0;
  } else {
    const long long per_e = static_cast<long long>(p.num_m_blocks) * p.num_n_blocks;
    ti.e = static_cast<int>(t / per_e);
    const int r = static_cast<int>(t % per_e);
    const int band_tiles = p.band * p.num_m_blocks;
    const int b = r / band_tiles, rr = r % band_tiles;
    const int nb0 = b * p.band;
    const int w = min(p.band, p.num_n_blocks - nb0);
    ti.mb = rr / w;
    ti.nb = nb0 + rr % w;
    if (p.dense) {
      ti.a_row = ti.e * p.a_expert_rows;
      ti.b_row = ti.e * p.b_expert_rows;
      ti.nkb = p.dense_kblocks;
    } else {
      const int r0 = __ldg(p.pad_offsets + ti.e), r1 = __ldg(p.pad_offsets + ti.e + 1);
      ti.a_row = r0;
      ti.b_row = r0;
      ti.nkb = (r1 - r0) / kBK;
    }
    ti.valid = true;
  }
  return ti;
}

template <typename OutT>
__device__ __forceinline__ void epilogue_store8(const KParams& p, const float (&acc)[8], OutT* c_row, OutT* pre_row,
                                                const void* bias_row, int col) {
  float z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) z[i] = acc[i];
  if (bias_row != nullptr) {
    float b[8];
    if (p.bias_fp32)
      load8(reinterpret_cast<const float*>(bias_row) + col, b);
    else
      load8(reinterpret_cast<const __nv_bfloat16*>(bias_row) + col, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] += b[i];
  }
  if (p.accumulate) {
    float old[8];
    load8(c_row + col, old);
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] += old[i];
  }
  if (p.act != CSMOE_ACT_NONE || pre_row != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = round_as(z[i], static_cast<const OutT*>(nullptr));
    if (pre_row != nullptr) store8(pre_row + col, z);
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = act_apply(z[i], p.act);
  }
  store8(c_row + col, z);
}

// One output tile (this CTA's 128 rows x BN columns) from TMEM to global memory.  `t_row` = TMEM address of this warp's
// lane quadrant in the accumulator stage, `out_row` = global output row of this thread, `half` = which half of the
// tile's columns this warp owns.
template <int MODE, int BN>
__device__ __forceinline__ void epilogue_tile(const KParams& p, const Tile& ti, uint32_t t_row, bool has_acc,
                                              long long out_row, int half) {
  bool row_ok = true;
  long long c_off = 0;
  if (MODE == CSMOE_GEMM_REDUCE) {
    row_ok = out_row < p.m_valid;
    c_off = static_cast<long long>(ti.e) * p.c_expert_stride;
  }
  const void* bias_row = nullptr;
  if (p.bias != nullptr) {
    const long long boff = static_cast<long long>(ti.e) * p.n;
    bias_row = p.bias_fp32 ? static_cast<const void*>(reinterpret_cast<const float*>(p.bias) + boff)
                           : static_cast<const void*>(reinterpret_cast<const __nv_bfloat16*>(p.bias) + boff);
  }
  if (p.epi == kEpiPlain) {
    // Expert-parallel return: the row goes straight to its slot in the source rank's buffer (peer memory over NVLink).
    unsigned long long row_ptr = 0ull;
    if (MODE == CSMOE_GEMM_ROWS && p.c_rows != nullptr) {
      row_ptr = __ldg(p.c_rows + out_row);
      row_ok = row_ptr != 0ull;
    }
#pragma unroll 1
    for (int chunk = half * (BN / 64); chunk < (half + 1) * (BN / 64); ++chunk) {
      uint32_t v[32];
      if (has_acc) {
        ptx::tmem_ld_32x32b_x32(t_row + chunk * 32, v);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      const int col0 = ti.nb * BN + chunk * 32;
      if (row_ok && col0 < p.n) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = col0 + g * 8;
          if (col < p.n) {
            float a8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a8[i] = __uint_as_float(v[g * 8 + i]);
            if (p.c_fp32) {
              float* c_row = row_ptr ? reinterpret_cast<float*>(row_ptr)
                                     : reinterpret_cast<float*>(p.c) + c_off + out_row * p.ldc;
              float* pre_row = p.preact ? reinterpret_cast<float*>(p.preact) + out_row * p.ldpre : nullptr;
              epilogue_store8<float>(p, a8, c_row, pre_row, bias_row, col);
            } else {
              __nv_bfloat16* c_row = row_ptr ? reinterpret_cast<__nv_bfloat16*>(row_ptr)
                                             : reinterpret_cast<__nv_bfloat16*>(p.c) + c_off + out_row * p.ldc;
              __nv_bfloat16* pre_row =
                  p.preact ? reinterpret_cast<__nv_bfloat16*>(p.preact) + out_row * p.ldpre : nullptr;
              epilogue_store8<__nv_bfloat16>(p, a8, c_row, pre_row, bias_row, col);
            }
          }
        }
      }
    }
  } else if (p.epi == kEpiGluFwd) {
    // TMEM columns [0,128) hold the gate and [128,256) the up projection of output columns nb*128 .. +128.
    // z = (gate | up) is stored for the backward pass, h = up * silu(gate) feeds the down projection
    // (Phi3MLP; each intermediate rounded to bf16 like the eager reference).
    if constexpr (BN == 256) {
      __nv_bfloat16* h_row = reinterpret_cast<__nv_bfloat16*>(p.c) + out_row * p.ldc;
      __nv_bfloat16* z_row = reinterpret_cast<__nv_bfloat16*>(p.preact) + out_row * p.ldpre;
#pragma unroll 1
      for (int chunk = half * 2; chunk < half * 2 + 2; ++chunk) {
        uint32_t vg[32], vu[32];
        ptx::tmem_ld_32x32b_x32(t_row + chunk * 32, vg);
        ptx::tmem_ld_32x32b_x32(t_row + 128 + chunk * 32, vu);
        ptx::tmem_ld_wait();
        const int col0 = ti.nb * 128 + chunk * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = col0 + g * 8;
          if (col < p.glu_f) {
            float zg[8], zu[8], h[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              zg[i] = bf16_round(__uint_as_float(vg[g * 8 + i]));
              zu[i] = bf16_round(__uint_as_float(vu[g * 8 + i]));
              h[i] = zu[i] * bf16_round(act_apply(zg[i], CSMOE_ACT_SILU));
            }
            store8(z_row + col, zg);
            store8(z_row + p.glu_f + col, zu);
            store8(h_row + col, h);
          }
        }
      }
    }
  } else {
    // Backward epilogues: the accumulator is dh; multiply by the activation derivative at the saved z.
    const __nv_bfloat16* z_row = reinterpret_cast<const __nv_bfloat16*>(p.aux) + out_row * p.ldaux;
    __nv_bfloat16* c_row = reinterpret_cast<__nv_bfloat16*>(p.c) + c_off + out_row * p.ldc;
    const bool glu = p.epi == kEpiGluBwd;
#pragma unroll 1
    for (int chunk = half * (BN / 64); chunk < (half + 1) * (BN / 64); ++chunk) {
      const int col0 = ti.nb * BN + chunk * 32;
      float z0[4][8], z1[4][8];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = col0 + g * 8;
        if (row_ok && col < p.n) {
          load8(z_row + col, z0[g]);
          if (glu) load8(z_row + p.glu_f + col, z1[g]);
        }
      }
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(t_row + chunk * 32, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = col0 + g * 8;
        if (row_ok && col < p.n) {
          float d0[8], d1[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float dh = bf16_round(__uint_as_float(v[g * 8 + i]));
            if (glu) {
              const float sg = bf16_round(act_apply(z0[g][i], CSMOE_ACT_SILU));
              d1[i] = dh * sg;                                                     // d up
              d0[i] = bf16_round(dh * z1[g][i]) * act_grad(z0[g][i], CSMOE_ACT_SILU);  // d gate
            } else {
              d0[i] = dh * act_grad(z0[g][i], p.act);
            }
          }
          store8(c_row + col, d0);
          if (glu) store8(c_row + p.glu_f + col, d1);
        }
      }
    }
  }
}

template <int MODE, bool B_MN, int BN>
__global__ void __launch_bounds__(kThreads, 1)
grouped_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const KParams p) {
  constexpr bool kAMn = (MODE == CSMOE_GEMM_REDUCE);
  constexpr bool kBMn = kAMn || B_MN;
  constexpr int kBBytes = BN * kBK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kStages = (BN == 256) ? 4 : 6;
  constexpr uint32_t kTmemCols = kAccStages * BN;
  constexpr uint32_t kIdesc = ptx::make_idesc_bf16(kBM, BN, kAMn, kBMn);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + kAccStages + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 2 * kAccStages);
  // generic pointer to the TMEM-address slot (same location as tmem_slot)
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < kAccStages; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const Tile ti = decode_tile<MODE>(p, t);
        if (!ti.valid) continue;
        for (int kb = 0; kb < ti.nkb; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t fb = full_bar(stage);
          ptx::mbar_arrive_expect_tx(fb, kStageBytes);
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint32_t sb = sa + kABytes;
          if (MODE == CSMOE_GEMM_ROWS) {
            ptx::tma_load_2d(sa, &tma_a, fb, kb * kBK, ti.a_row);
            if (!B_MN) {
              if (BN == 256 && p.epi == kEpiGluFwd) {
                // gate rows [nb*128, +128) then up rows [F + nb*128, +128): one output tile holds both halves
                ptx::tma_load_3d(sb, &tma_b, fb, kb * kBK, ti.nb * 128, ti.e);
                ptx::tma_load_3d(sb + 128 * kBK * 2, &tma_b, fb, kb * kBK, p.glu_f + ti.nb * 128, ti.e);
              } else {
                ptx::tma_load_3d(sb, &tma_b, fb, kb * kBK, ti.nb * BN, ti.e);
              }
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                ptx::tma_load_3d(sb + j * kSubTileBytes, &tma_b, fb, ti.nb * BN + j * 64, kb * kBK, ti.e);
            }
          } else {
#pragma unroll
            for (int j = 0; j < kBM / 64; ++j)
              ptx::tma_load_2d(sa + j * kSubTileBytes, &tma_a, fb, ti.mb * kBM + j * 64, ti.a_row + kb * kBK);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              ptx::tma_load_2d(sb + j * kSubTileBytes, &tma_b, fb, ti.nb * BN + j * 64, ti.b_row + kb * kBK);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const Tile ti = decode_tile<MODE>(p, t);
        if (!ti.valid || ti.nkb == 0) continue;
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < ti.nkb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // K-major: 16 k-elements = 32 bytes inside the 128B swizzle row; 8-row groups are 1024 B apart.
            // MN-major: 16 k-rows = 2048 bytes; 64-wide mn blocks are kSubTileBytes apart, 8-k groups 1024 B apart.
            const uint64_t adesc = kAMn ? ptx::make_smem_desc_sw128(sa + k * 2048, kSubTileBytes, 1024)
                                        : ptx::make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t bdesc = kBMn ? ptx::make_smem_desc_sw128(sb + k * 2048, kSubTileBytes, 1024)
                                        : ptx::make_smem_desc_sw128(sb + k * 32, 16, 1024);
            ptx::umma_f16(d_tmem, adesc, bdesc, kIdesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar(stage));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        ptx::umma_commit(tfull_bar(acc));
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================================================== epilogue (8 warps: 128 TMEM lanes x 2 column halves)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quad * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const Tile ti = decode_tile<MODE>(p, t);
      if (!ti.valid) continue;
      const bool has_acc = ti.nkb > 0;
      if (has_acc) {
        ptx::mbar_wait(tfull_bar(acc), acc_phase);
        ptx::tc_fence_after();
      }
      const long long out_row = static_cast<long long>(ti.mb) * kBM + row_in_tile;
      const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
      epilogue_tile<MODE, BN>(p, ti, t_row, has_acc, out_row, half);
      if (has_acc) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ CTA-pair kernel
// Same roles and pipeline, but two CTAs of a cluster (one TPC) cooperate on a 256 x 256 tile with
// tcgen05.mma.cta_group::2: each CTA stages its own 128 rows of A and HALF of B (128 of the 256 tile columns), the
// leader CTA issues the MMAs for both, and each CTA's TMEM receives its 128 rows x 256 columns.  Per-SM shared-memory
// traffic per MMA drops from A+B to A+B/2, which is what keeps the tensor pipe fed (GEMM section).
//   full[s]   lives in the leader: 1 arrival (its expect_tx) + the bytes of BOTH CTAs' TMA loads
//   empty[s]  one per CTA, signalled by the leader's tcgen05.commit multicast
//   tfull[a]  one per CTA, same multicast commit after the last k-block of a tile
//   tempty[a] lives in the leader: 2 x kEpiWarps arrivals (the peer's epilogue warps arrive remotely)
constexpr int kPairStages = 6;
constexpr int kPairBBytes = 128 * kBK * 2;
constexpr int kPairStageBytes = kABytes + kPairBBytes;  // 32 KiB per CTA per stage

template <int MODE>
__device__ __forceinline__ Tile decode_tile_pair(const KParams& p, long long t, int rank) {
  Tile ti;
  const int num_m2 = p.num_m_pairs;
  if (MODE == CSMOE_GEMM_ROWS) {
    const long long band_tiles = static_cast<long long>(p.band) * num_m2;
    const int b = static_cast<int>(t / band_tiles);
    const int r = static_cast<int>(t % band_tiles);
    const int nb0 = b * p.band;
    const int w = min(p.band, p.num_n_blocks - nb0);
    ti.mb = r / w;
    ti.nb = nb0 + r % w;
    if (p.dense) {
      const int dm2 = p.dense_mblocks / 2;
      ti.e = ti.mb / dm2;
      ti.a_row = (ti.mb % dm2) * 256 + rank * kBM + ti.e * p.a_expert_rows;
    } else {
      ti.e = __ldg(p.tile_expert + 2 * ti.mb);
      ti.a_row = ti.mb * 256 + rank * kBM;
    }
    ti.b_row = 0;
    ti.nkb = p.num_kb;
    ti.valid = ti.e >= 0;
  } else {
    const long long per_e = static_cast<long long>(num_m2) * p.num_n_blocks;
    ti.e = static_cast<int>(t / per_e);
    const int r = static_cast<int>(t % per_e);
    const int band_tiles = p.band * num_m2;
    const int b = r / band_tiles, rr = r % band_tiles;
    const int nb0 = b * p.band;
    const int w = min(p.band, p.num_n_blocks - nb0);
    ti.mb = rr / w;
    ti.nb = nb0 + rr % w;
    if (p.dense) {
      ti.a_row = ti.e * p.a_expert_rows;
      ti.b_row = ti.e * p.b_expert_rows;
      ti.nkb = p.dense_kblocks;
    } else {
      const int r0 = __ldg(p.pad_offsets + ti.e), r1 = __ldg(p.pad_offsets + ti.e + 1);
      ti.a_row = r0;
      ti.b_row = r0;
      ti.nkb = (r1 - r0) / kBK;
    }
    ti.valid = true;
  }
  return ti;
}

template <int MODE, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
grouped_gemm_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                         const KParams p) {
  constexpr int BN = 256;
  constexpr bool kAMn = (MODE == CSMOE_GEMM_REDUCE);
  constexpr bool kBMn = kAMn || B_MN;
  constexpr uint32_t kTmemCols = kAccStages * BN;
  constexpr uint32_t kIdesc = ptx::make_idesc_bf16(256, BN, kAMn, kBMn);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kPairStages * kPairStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kPairStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kPairStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kPairStages + kAccStages + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kPairStages + 2 * kAccStages);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(ptx::cluster_ctarank());
  const bool leader = rank == 0;
  const long long cluster_id = blockIdx.x >> 1;
  const long long num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kPairStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < kAccStages; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), 2 * kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_cg2(tmem_slot, kTmemCols);
    ptx::tmem_relinquish_cg2();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long t = cluster_id; t < p.total_tiles; t += num_clusters) {
        const Tile ti = decode_tile_pair<MODE>(p, t, rank);
        if (!ti.valid) continue;
        for (int kb = 0; kb < ti.nkb; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          if (leader) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * kPairStageBytes);
          const uint32_t fb = ptx::mapa(full_bar(stage), 0);  // the leader's barrier collects both CTAs' bytes
          const uint32_t sa = smem_base + stage * kPairStageBytes;
          const uint32_t sb = sa + kABytes;
          if (MODE == CSMOE_GEMM_ROWS) {
            ptx::tma_load_2d_cg2(sa, &tma_a, fb, kb * kBK, ti.a_row);
            if (!B_MN) {
              const int brow = p.epi == kEpiGluFwd ? (rank == 0 ? ti.nb * 128 : p.glu_f + ti.nb * 128)
                                                   : ti.nb * BN + rank * 128;
              ptx::tma_load_3d_cg2(sb, &tma_b, fb, kb * kBK, brow, ti.e);
            } else {
#pragma unroll
              for (int j = 0; j < 2; ++j)
                ptx::tma_load_3d_cg2(sb + j * kSubTileBytes, &tma_b, fb, ti.nb * BN + rank * 128 + j * 64, kb * kBK, ti.e);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              ptx::tma_load_2d_cg2(sa + j * kSubTileBytes, &tma_a, fb, ti.mb * 256 + rank * 128 + j * 64,
                                   ti.a_row + kb * kBK);
#pragma unroll
            for (int j = 0; j < 2; ++j)
              ptx::tma_load_2d_cg2(sb + j * kSubTileBytes, &tma_b, fb, ti.nb * BN + rank * 128 + j * 64,
                                   ti.b_row + kb * kBK);
          }
          if (++stage == kPairStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (long long t = cluster_id; t < p.total_tiles; t += num_clusters) {
        const Tile ti = decode_tile_pair<MODE>(p, t, rank);
        if (!ti.valid || ti.nkb == 0) continue;
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < ti.nkb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * kPairStageBytes;
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t adesc = kAMn ? ptx::make_smem_desc_sw128(sa + k * 2048, kSubTileBytes, 1024)
                                        : ptx::make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t bdesc = kBMn ? ptx::make_smem_desc_sw128(sb + k * 2048, kSubTileBytes, 1024)
                                        : ptx::make_smem_desc_sw128(sb + k * 32, 16, 1024);
            ptx::umma_f16_cg2(d_tmem, adesc, bdesc, kIdesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit_cg2_mc(empty_bar(stage), 0x3);
          if (++stage == kPairStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        ptx::umma_commit_cg2_mc(tfull_bar(acc), 0x3);
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================================================== epilogue (both CTAs; 8 warps each)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quad * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    for (long long t = cluster_id; t < p.total_tiles; t += num_clusters) {
      const Tile ti = decode_tile_pair<MODE>(p, t, rank);
      if (!ti.valid) continue;
      const bool has_acc = ti.nkb > 0;
      if (has_acc) {
        ptx::mbar_wait(tfull_bar(acc), acc_phase);
        ptx::tc_fence_after();
      }
      const long long out_row = static_cast<long long>(ti.mb) * 256 + rank * kBM + row_in_tile;
      const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
      epilogue_tile<MODE, BN>(p, ti, t_row, has_acc, out_row, half);
      if (has_acc) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader)
            ptx::mbar_arrive(tempty_bar(acc));
          else
            ptx::mbar_arrive_remote(ptx::mapa(tempty_bar(acc), 0));
        }
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();  // the peer's smem / barriers must outlive the leader's last MMA and commit
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fn() {
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

int encode_bf16_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box) {
  EncodeFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return CSMOE_ERR_DRIVER;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), dims,
                  strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu,%llu stride0 %llu box %u,%u)", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              (unsigned long long)strides_bytes[0], box[0], box[1]);
    return CSMOE_ERR_DRIVER;
  }
  return CSMOE_OK;
}

template <int MODE, bool B_MN, int BN>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const KParams& kp, int grid, cudaStream_t stream) {
  constexpr int kStages = (BN == 256) ? 4 : 6;
  constexpr int kSmem = kStages * (kABytes + BN * kBK * 2) + 1024 + 256;
  auto kern = grouped_gemm_kernel<MODE, B_MN, BN>;
  static bool configured = false;  // per instantiation; benign race (idempotent attribute set)
  if (!configured) {
    CSMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  kern<<<grid, kThreads, kSmem, stream>>>(ma, mb, kp);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

template <int MODE, bool B_MN>
int launch_pair(const CUtensorMap& ma, const CUtensorMap& mb, const KParams& kp, int clusters, cudaStream_t stream) {
  constexpr int kSmem = kPairStages * kPairStageBytes + 1024 + 256;
  auto kern = grouped_gemm_pair_kernel<MODE, B_MN>;
  static bool configured = false;
  if (!configured) {
    CSMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  kern<<<2 * clusters, kThreads, kSmem, stream>>>(ma, mb, kp);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

// CSMOE_GEMM_PAIR=0 disables the CTA-pair kernel (A/B comparisons, bring-up)
bool pair_enabled() {
  static const bool on = []() {
    const char* v = getenv("CSMOE_GEMM_PAIR");
    return v == nullptr || v[0] != '0';
  }();
  return on;
}

}  // namespace
}  // namespace csmoe

using namespace csmoe;

extern "C" int csmoe_grouped_gemm(const csmoe_gemm_args* a, void* stream_) {
  CSMOE_CHECK_ARG(a != nullptr, "csmoe_grouped_gemm: args is NULL");
  CSMOE_CHECK_ARG(a->a && a->b && (a->c || a->c_rows), "csmoe_grouped_gemm: a/b/c must be non-NULL");
  CSMOE_CHECK_ARG(a->c_rows == nullptr || (a->mode == CSMOE_GEMM_ROWS && a->act == CSMOE_ACT_NONE &&
                                           a->act_bwd == CSMOE_ACT_NONE && a->preact == nullptr && !a->accumulate),
                  "csmoe_grouped_gemm: c_rows (per-row destinations) needs ROWS mode and the plain (bias-only) epilogue");
  CSMOE_CHECK_ARG(a->mode == CSMOE_GEMM_ROWS || a->mode == CSMOE_GEMM_REDUCE, "csmoe_grouped_gemm: bad mode %d", a->mode);
  CSMOE_CHECK_ARG(a->num_experts >= 1, "csmoe_grouped_gemm: num_experts must be >= 1");
  CSMOE_CHECK_ARG(a->m > 0 && a->n > 0 && a->k > 0, "csmoe_grouped_gemm: m, n, k must be positive");
  CSMOE_CHECK_ARG(a->n % 8 == 0, "csmoe_grouped_gemm: n (%lld) must be a multiple of 8", (long long)a->n);
  CSMOE_CHECK_ARG(a->lda % 8 == 0 && a->ldb % 8 == 0, "csmoe_grouped_gemm: lda/ldb must be multiples of 8 elements");
  CSMOE_CHECK_ARG(a->ldc % (a->c_dtype == CSMOE_F32 ? 4 : 8) == 0, "csmoe_grouped_gemm: ldc alignment");
  CSMOE_CHECK_ARG((reinterpret_cast<uintptr_t>(a->a) | reinterpret_cast<uintptr_t>(a->b) |
                   reinterpret_cast<uintptr_t>(a->c)) % 16 == 0,
                  "csmoe_grouped_gemm: a/b/c must be 16-byte aligned");
  CSMOE_CHECK_ARG(a->c_dtype == CSMOE_F32 || a->c_dtype == CSMOE_BF16, "csmoe_grouped_gemm: bad c_dtype");
  const bool glu_fwd = a->act == CSMOE_ACT_SILU_GLU;
  const bool act_bwd = a->act_bwd != CSMOE_ACT_NONE;
  if (glu_fwd) {
    CSMOE_CHECK_ARG(a->mode == CSMOE_GEMM_ROWS && a->b_layout == 0 && a->preact && a->c_dtype == CSMOE_BF16 &&
                        a->n % 2 == 0 && (a->n / 2) % 8 == 0 && a->bias == nullptr && !act_bwd,
                    "csmoe_grouped_gemm: fused SILU_GLU needs ROWS mode, [2F,k] weights, bf16 C [m,F], preact [m,2F], no bias");
  }
  if (act_bwd) {
    CSMOE_CHECK_ARG(a->mode == CSMOE_GEMM_ROWS && a->aux && a->c_dtype == CSMOE_BF16 && a->act == CSMOE_ACT_NONE &&
                        a->bias == nullptr && a->preact == nullptr && a->ldaux % 8 == 0,
                    "csmoe_grouped_gemm: act_bwd epilogue needs ROWS mode, bf16 C, aux = saved pre-activation, no bias/act");
  }
  if (a->dense) {
    CSMOE_CHECK_ARG(a->dense_rows > 0 && a->dense_rows % kBM == 0, "csmoe_grouped_gemm: dense_rows must be a multiple of 128");
  } else if (a->mode == CSMOE_GEMM_ROWS) {
    CSMOE_CHECK_ARG(a->tile_expert != nullptr, "csmoe_grouped_gemm: ROWS mode needs tile_expert");
    CSMOE_CHECK_ARG(a->m % kBM == 0, "csmoe_grouped_gemm: ROWS mode m must be a multiple of 128");
  } else {
    CSMOE_CHECK_ARG(a->pad_offsets != nullptr, "csmoe_grouped_gemm: REDUCE mode needs pad_offsets");
  }
  if (a->accumulate) CSMOE_CHECK_ARG(a->c_dtype == CSMOE_F32, "csmoe_grouped_gemm: accumulate needs an fp32 C");

  cudaStream_t stream = as_stream(stream_);
  const int E = a->num_experts;
  // n as seen by the tile grid: the fused GLU forward produces F output columns from 2F weight rows
  const long long n_grid = glu_fwd ? a->n / 2 : a->n;
  const bool big_n = glu_fwd || a->n > 128;
  const int BN = big_n ? 256 : 128;
  // CTA-pair (256 x 256 tile) eligibility: wide outputs, and row tiles that never straddle experts in 256-row units
  bool pair = pair_enabled() && big_n && a->n >= 256;
  if (a->mode == CSMOE_GEMM_ROWS) {
    if (a->dense)
      pair = pair && (a->dense_rows % 256 == 0) && (a->a_expert_rows % 256 == 0);
    else
      pair = pair && a->row_tile >= 256 && (a->m % 256 == 0);
  } else {
    pair = pair && a->m >= 256;
  }

  KParams kp{};
  kp.n = static_cast<int>(a->n);
  kp.num_experts = E;
  kp.dense = a->dense;
  kp.dense_mblocks = a->dense ? static_cast<int>(a->dense_rows / kBM) : 0;
  kp.dense_kblocks = a->dense ? static_cast<int>(a->dense_rows / kBK) : 0;
  kp.a_expert_rows = static_cast<int>(a->a_expert_rows);
  kp.act = a->act;
  kp.c_fp32 = a->c_dtype == CSMOE_F32;
  kp.bias_fp32 = a->bias_dtype == CSMOE_F32;
  kp.accumulate = a->accumulate;
  kp.tile_expert = a->tile_expert;
  kp.pad_offsets = a->pad_offsets;
  kp.c = a->c;
  kp.preact = a->preact;
  kp.bias = a->bias;
  kp.ldc = a->ldc;
  kp.ldpre = a->ldpre;
  kp.c_expert_stride = a->c_expert_stride;
  kp.num_n_blocks = glu_fwd ? static_cast<int>((n_grid + 127) / 128) : static_cast<int>((a->n + BN - 1) / BN);
  kp.band = 8;
  kp.aux = a->aux;
  kp.ldaux = a->ldaux;
  kp.c_rows = reinterpret_cast<const unsigned long long*>(a->c_rows);
  if (glu_fwd) {
    kp.epi = kEpiGluFwd;
    kp.glu_f = static_cast<int>(a->n / 2);
    kp.act = CSMOE_ACT_NONE;
  } else if (act_bwd) {
    kp.epi = a->act_bwd == CSMOE_ACT_SILU_GLU ? kEpiGluBwd : kEpiActBwd;
    kp.glu_f = static_cast<int>(a->n);
    kp.act = a->act_bwd;
  } else {
    kp.epi = kEpiPlain;
  }

  CUtensorMap ma, mb;
  int rc;
  if (a->mode == CSMOE_GEMM_ROWS) {
    const long long a_rows = a->dense ? (a->a_expert_rows ? (long long)E * a->a_expert_rows : a->dense_rows) : a->m;
    kp.num_m_blocks = a->dense ? E * kp.dense_mblocks : static_cast<int>(a->m / kBM);
    kp.num_kb = static_cast<int>((a->k + kBK - 1) / kBK);
    kp.total_tiles = static_cast<long long>(kp.num_m_blocks) * kp.num_n_blocks;
    {
      cuuint64_t dims[2] = {(cuuint64_t)a->k, (cuuint64_t)a_rows};
      cuuint64_t str[1] = {(cuuint64_t)a->lda * 2};
      cuuint32_t box[2] = {kBK, kBM};
      if ((rc = encode_bf16_map(&ma, a->a, 2, dims, str, box)) != CSMOE_OK) return rc;
    }
    if (a->b_layout == 0) {
      cuuint64_t dims[3] = {(cuuint64_t)a->k, (cuuint64_t)a->n, (cuuint64_t)E};
      cuuint64_t str[2] = {(cuuint64_t)a->ldb * 2, (cuuint64_t)a->b_expert_stride * 2};
      if (E == 1) str[1] = (cuuint64_t)a->ldb * 2 * a->n;
      cuuint32_t box[3] = {kBK, (cuuint32_t)((glu_fwd || pair) ? 128 : BN), 1};
      if ((rc = encode_bf16_map(&mb, a->b, 3, dims, str, box)) != CSMOE_OK) return rc;
    } else {
      cuuint64_t dims[3] = {(cuuint64_t)a->n, (cuuint64_t)a->k, (cuuint64_t)E};
      cuuint64_t str[2] = {(cuuint64_t)a->ldb * 2, (cuuint64_t)a->b_expert_stride * 2};
      if (E == 1) str[1] = (cuuint64_t)a->ldb * 2 * a->k;
      cuuint32_t box[3] = {64, kBK, 1};
      if ((rc = encode_bf16_map(&mb, a->b, 3, dims, str, box)) != CSMOE_OK) return rc;
    }
  } else {
    kp.m_valid = static_cast<int>(a->m);
    kp.num_m_blocks = static_cast<int>((a->m + kBM - 1) / kBM);
    kp.b_expert_rows = static_cast<int>(a->b_expert_stride);
    kp.total_tiles = static_cast<long long>(E) * kp.num_m_blocks * kp.num_n_blocks;
    const long long a_rows = a->dense ? (a->a_expert_rows ? (long long)E * a->a_expert_rows : a->dense_rows) : a->k;
    const long long b_rows = a->dense ? (a->b_expert_stride ? (long long)E * a->b_expert_stride : a->dense_rows) : a->k;
    {
      cuuint64_t dims[2] = {(cuuint64_t)a->m, (cuuint64_t)a_rows};
      cuuint64_t str[1] = {(cuuint64_t)a->lda * 2};
      cuuint32_t box[2] = {64, kBK};
      if ((rc = encode_bf16_map(&ma, a->a, 2, dims, str, box)) != CSMOE_OK) return rc;
    }
    {
      cuuint64_t dims[2] = {(cuuint64_t)a->n, (cuuint64_t)b_rows};
      cuuint64_t str[1] = {(cuuint64_t)a->ldb * 2};
      cuuint32_t box[2] = {64, kBK};
      if ((rc = encode_bf16_map(&mb, a->b, 2, dims, str, box)) != CSMOE_OK) return rc;
    }
  }
  if (kp.total_tiles == 0) return CSMOE_OK;
  if (pair) {
    kp.num_m_pairs = a->mode == CSMOE_GEMM_ROWS ? kp.num_m_blocks / 2 : static_cast<int>((a->m + 255) / 256);
    const long long per = static_cast<long long>(kp.num_m_pairs) * kp.num_n_blocks;
    kp.total_tiles = a->mode == CSMOE_GEMM_ROWS ? per : per * E;
    int clusters = num_sms() / 2;
    if (clusters <= 0) return CSMOE_ERR_CUDA;
    if (a->max_ctas > 1 && a->max_ctas / 2 < clusters) clusters = a->max_ctas / 2;
    if (kp.total_tiles < clusters) clusters = static_cast<int>(kp.total_tiles);
    if (a->mode == CSMOE_GEMM_ROWS)
      return a->b_layout == 0 ? launch_pair<CSMOE_GEMM_ROWS, false>(ma, mb, kp, clusters, stream)
                              : launch_pair<CSMOE_GEMM_ROWS, true>(ma, mb, kp, clusters, stream);
    return launch_pair<CSMOE_GEMM_REDUCE, true>(ma, mb, kp, clusters, stream);
  }
  int grid = num_sms();
  if (grid <= 0) return CSMOE_ERR_CUDA;
  if (a->max_ctas > 0 && a->max_ctas < grid) grid = a->max_ctas;
  if (kp.total_tiles < grid) grid = static_cast<int>(kp.total_tiles);

  if (a->mode == CSMOE_GEMM_ROWS) {
    if (a->b_layout == 0)
      return big_n ? launch<CSMOE_GEMM_ROWS, false, 256>(ma, mb, kp, grid, stream)
                   : launch<CSMOE_GEMM_ROWS, false, 128>(ma, mb, kp, grid, stream);
    return big_n ? launch<CSMOE_GEMM_ROWS, true, 256>(ma, mb, kp, grid, stream)
                 : launch<CSMOE_GEMM_ROWS, true, 128>(ma, mb, kp, grid, stream);
  }
  return big_n ? launch<CSMOE_GEMM_REDUCE, true, 256>(ma, mb, kp, grid, stream)
               : launch<CSMOE_GEMM_REDUCE, true, 128>(ma, mb, kp, grid, stream);
}
