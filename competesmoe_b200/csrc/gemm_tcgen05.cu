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

// The quad-cluster variant (clusters of two CTA pairs with the A tile multicast; measured slower, profiles/r01n_quad_cluster.md)
// is kept in the source for the record but compiled only with -DCSMOE_BUILD_QUAD=1: three fewer cubins in the default build.
#ifndef CSMOE_BUILD_QUAD
#define CSMOE_BUILD_QUAD 0
#endif

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
  int bias_after_round;  // 1: C = bf16(acc) + bias (an fp32 bias added to a bf16 op result: the pretrain plugin, moe.py:400-401)
  int accumulate;
  int glu_f;          // GLU epilogues: F (forward: n == 2F and tiles interleave gate|up; backward: n == F)
  int epi;            // kEpiPlain / kEpiGluFwd / kEpiActBwd / kEpiGluBwd
  int band;           // n-blocks per rasterisation band
  int raster_m;       // pair / wide ROWS launches: 1 = bands of m-blocks with m fastest (see decode_tile_pair);
                      // 2 = the same with the bands aligned to the experts' row ranges (needs pad_offsets)
  int band_cap;       // raster_m == 2: largest band (m-blocks) an expert's row range is swept in
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
  float* rowsum;      // plain ROWS epilogue: rowsum[row, col/64] = sum over the 64-column group of softplus(stored value)
  int rowsum_ld;      // (the competition step's neural-response score, reduced in the epilogue); groups per row
  int rowsum_round;   // 1 = round every softplus to bf16 first (eager bf16 reference arithmetic)
  int kcat;           // ROWS + dense: C[rows, n] = sum_e A[e*a_expert_rows + rows, k] . B[e]  (the k loop runs over experts too)
  int dbg_mode;       // tuning experiments (CSMOE_GEMM_DBG): 1 = no TMA loads after the first pipeline fill, 2 = no MMAs
  int direct_epi;     // 1 = register-direct (row per thread) epilogue stores instead of the staged, coalesced ones
  int tma_epi;        // 1 = 32 x 32 groups leave through TMA stores (tensor maps tma_c / tma_p); 2 = backward epilogue of the
                      // pair kernel: the saved pre-activation arrives by TMA loads (tma_z) as well
  int stages;         // pair kernel: operand stages in use (6, or 5 when the epilogue needs 8 KiB of staging per warp)
  int stg_bytes;      // pair kernel: staging bytes per epilogue warp (4096 / 8192)
  unsigned long long* stats;  // debug (CSMOE_GEMM_STATS=1): per CTA {producer wait, mma wait full, mma wait tempty, epilogue, total, tiles}
};

__device__ __forceinline__ void timed_wait(uint32_t bar, uint32_t parity, bool on, unsigned long long& acc) {
  if (!on) {
    ptx::mbar_wait(bar, parity);
    return;
  }
  const long long t0 = clock64();
  ptx::mbar_wait(bar, parity);
  acc += static_cast<unsigned long long>(clock64() - t0);
}

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
    if (p.kcat) {
      ti.e = 0;
      ti.a_row = ti.mb * kBM;
    } else if (p.dense) {
      ti.e = ti.mb / p.dense_mblocks;
      ti.a_row = (ti.mb % p.dense_mblocks) * kBM + ti.e * p.a_expert_rows;
    } else {
      ti.e = __ldg(p.tile_expert + ti.mb);
      ti.a_row = ti.mb * kBM;
    }
    ti.b_row = 0;
    ti.nkb = p.kcat ? p.num_kb * p.num_experts : p.num_kb;
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

__device__ __forceinline__ float softplus_fast(float x) { return fmaxf(x, 0.f) + __logf(1.f + __expf(-fabsf(x))); }

template <typename OutT>
__device__ __forceinline__ void epilogue_store8(const KParams& p, const float (&acc)[8], OutT* c_row, OutT* pre_row,
                                                const void* bias_row, int col, float* rs = nullptr) {
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
    for (int i = 0; i < 8; ++i) z[i] = (p.bias_after_round ? bf16_round(z[i]) : z[i]) + b[i];
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
    act_apply_vec<8>(z, p.act, !p.c_fp32);
  }
  store8(c_row + col, z);
  if (rs != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float sp = softplus_fast(round_as(z[i], static_cast<const OutT*>(nullptr)));   // of the value as stored
      *rs += p.rowsum_round ? bf16_round(sp) : sp;
    }
  }
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
    float rs = 0.f;
    float* rs_ptr = (MODE == CSMOE_GEMM_ROWS && p.rowsum != nullptr) ? &rs : nullptr;
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
              epilogue_store8<float>(p, a8, c_row, pre_row, bias_row, col, rs_ptr);
            } else {
              __nv_bfloat16* c_row = row_ptr ? reinterpret_cast<__nv_bfloat16*>(row_ptr)
                                             : reinterpret_cast<__nv_bfloat16*>(p.c) + c_off + out_row * p.ldc;
              __nv_bfloat16* pre_row =
                  p.preact ? reinterpret_cast<__nv_bfloat16*>(p.preact) + out_row * p.ldpre : nullptr;
              epilogue_store8<__nv_bfloat16>(p, a8, c_row, pre_row, bias_row, col, rs_ptr);
            }
          }
        }
      }
      // every second 32-column chunk closes a 64-column group of the row-sum output
      if (rs_ptr != nullptr && (chunk & 1)) {
        if (col0 - 32 < p.n) p.rowsum[out_row * p.rowsum_ld + (col0 >> 6)] = rs;
        rs = 0.f;
      }
    }
  } else if (p.epi == kEpiGluFwd) {
    // TMEM columns [0,128) hold the gate and [128,256) the up projection of output columns nb*128 .. +128.
    // z = (gate | up) is stored for the backward pass, h = up * silu(gate) feeds the down projection
    // (Phi3MLP; each intermediate rounded to bf16 like the eager reference).
    if constexpr (BN >= 256) {
      constexpr int kGate = BN / 2;          // TMEM columns [0, kGate) = gate, [kGate, BN) = up, of kGate output columns
      __nv_bfloat16* h_row = reinterpret_cast<__nv_bfloat16*>(p.c) + out_row * p.ldc;
      __nv_bfloat16* z_row = reinterpret_cast<__nv_bfloat16*>(p.preact) + out_row * p.ldpre;
#pragma unroll 1
      for (int chunk = half * (kGate / 64); chunk < (half + 1) * (kGate / 64); ++chunk) {
        uint32_t vg[32], vu[32];
        ptx::tmem_ld_32x32b_x32(t_row + chunk * 32, vg);
        ptx::tmem_ld_32x32b_x32(t_row + kGate + chunk * 32, vu);
        ptx::tmem_ld_wait();
        const int col0 = ti.nb * kGate + chunk * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = col0 + g * 8;
          if (col < p.glu_f) {
            float zg[8], zu[8], h[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              zg[i] = bf16_round(__uint_as_float(vg[g * 8 + i]));
              zu[i] = bf16_round(__uint_as_float(vu[g * 8 + i]));
              h[i] = zu[i] * bf16_round(act_apply(zg[i], CSMOE_ACT_SILU, true));
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
              const float sg = bf16_round(act_apply(z0[g][i], CSMOE_ACT_SILU, true));
              d1[i] = dh * sg;                                                     // d up
              d0[i] = bf16_round(dh * z1[g][i]) * act_grad(z0[g][i], CSMOE_ACT_SILU, true);  // d gate
            } else {
              d0[i] = dh;
            }
          }
          if (!glu) act_grad_vec<8>(d0, z0[g], p.act, true);
          store8(c_row + col, d0);
          if (glu) store8(c_row + p.glu_f + col, d1);
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------ staged epilogue
// tcgen05.ld hands every thread one output ROW (lane = TMEM lane), so storing from registers makes each warp-level
// store touch 32 different rows, 16 bytes each: 32 partial-sector L2 writes per instruction.  The staged epilogue
// instead transposes 32-row x 32-column groups through a per-warp shared-memory tile (4 KiB, 16-byte chunks
// XOR-swizzled so both directions are bank-conflict free) and writes / reads global memory with 64- or 128-byte
// contiguous row segments, several rows per instruction.  The same tile stages the saved pre-activation z that the
// backward epilogues read.
constexpr int kStageTileBytes = 32 * 128;

// Row pitch P = 32 columns * element size: 64 (bf16) or 128 (fp32) bytes.
template <int P>
__device__ __forceinline__ uint32_t stg_addr(uint32_t stg, int row, int chunk) {
  if (P == 128) return stg + row * 128 + ((chunk ^ (row & 7)) << 4);
  return stg + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src) {
  const uint32_t lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v), src);
  const uint32_t hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), src);
  return (static_cast<unsigned long long>(hi) << 32) | lo;
}

// This thread's 32 values f[] (one row, 32 consecutive columns) -> global: row r of the group goes to byte address
// row_base(r) + col_byte; row_base is held by lane r (0 = row not stored).  `valid_cols` <= 32 columns are written.
template <bool FP32>
__device__ __forceinline__ void staged_store32(uint32_t stg, int lane, const float (&f)[32], unsigned long long my_row_base,
                                               long long col_byte, int valid_cols) {
  constexpr int P = FP32 ? 128 : 64;
  constexpr int kChunks = P / 16;          // 16-byte chunks per row
  constexpr int kRowsPerIt = 32 / kChunks;  // rows covered by one warp-wide 16-byte access
  if (FP32) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      sts128(stg_addr<P>(stg, lane, c), __float_as_uint(f[4 * c]), __float_as_uint(f[4 * c + 1]),
             __float_as_uint(f[4 * c + 2]), __float_as_uint(f[4 * c + 3]));
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      sts128(stg_addr<P>(stg, lane, c), pack_bf16(f[8 * c], f[8 * c + 1]), pack_bf16(f[8 * c + 2], f[8 * c + 3]),
             pack_bf16(f[8 * c + 4], f[8 * c + 5]), pack_bf16(f[8 * c + 6], f[8 * c + 7]));
  }
  __syncwarp();
  const int chunk = lane % kChunks;
  const bool col_ok = chunk * (FP32 ? 4 : 8) < valid_cols;
#pragma unroll
  for (int it = 0; it < 32 / kRowsPerIt; ++it) {
    const int r = it * kRowsPerIt + lane / kChunks;
    const unsigned long long base = shfl_u64(my_row_base, r);
    const uint4 v = lds128(stg_addr<P>(stg, r, chunk));
    if (base != 0ull && col_ok) *reinterpret_cast<uint4*>(base + col_byte + chunk * 16) = v;
  }
  __syncwarp();
}

// Saved bf16 pre-activations: 32 rows x 32 columns from global (row r at byte address row_base(r) + col_byte, held by
// lane r; 0 = row absent -> zeros) into this thread's z[32] (its own row).
__device__ __forceinline__ void staged_load32_bf16(uint32_t stg, int lane, float (&z)[32], unsigned long long my_row_base,
                                                   long long col_byte, int valid_cols) {
  constexpr int P = 64;
  const int chunk = lane & 3;
  const bool col_ok = chunk * 8 < valid_cols;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2);
    const unsigned long long base = shfl_u64(my_row_base, r);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (base != 0ull && col_ok) v = *reinterpret_cast<const uint4*>(base + col_byte + chunk * 16);
    sts128(stg_addr<P>(stg, r, chunk), v.x, v.y, v.z, v.w);
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 v = lds128(stg_addr<P>(stg, lane, c));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      z[8 * c + 2 * i] = t.x;
      z[8 * c + 2 * i + 1] = t.y;
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void tmem_ld32_f(uint32_t taddr, bool has_acc, float (&f)[32]) {
  if (has_acc) {
    uint32_t v[32];
    ptx::tmem_ld_32x32b_x32(taddr, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = 0.f;
  }
}

// One output tile of this warp: rows row0 .. row0+31 (lane = row), the half of the tile's BN columns given by `half`.
// `t_row` = TMEM address of the warp's lane quadrant in the accumulator, `stg` = the warp's staging tile.
template <int MODE, int BN>
__device__ __forceinline__ void epilogue_tile_staged(const KParams& p, const Tile& ti, uint32_t t_row, bool has_acc,
                                                     long long row0, int half, int lane, uint32_t stg) {
  const long long my_row = row0 + lane;
  bool row_ok = true;
  long long c_off = 0;
  if (MODE == CSMOE_GEMM_REDUCE) {
    row_ok = my_row < p.m_valid;
    c_off = static_cast<long long>(ti.e) * p.c_expert_stride;
  }
  const int esz = p.c_fp32 ? 4 : 2;
  unsigned long long my_c = 0ull;
  if (MODE == CSMOE_GEMM_ROWS && p.c_rows != nullptr)
    my_c = __ldg(p.c_rows + my_row);   // expert-parallel return: the row's slot in the source rank's buffer (0 = skip)
  else if (row_ok)
    my_c = reinterpret_cast<unsigned long long>(p.c) + static_cast<unsigned long long>((c_off + my_row * p.ldc) * esz);

  if (p.epi == kEpiPlain) {
    const unsigned long long my_pre =
        (p.preact != nullptr && row_ok) ? reinterpret_cast<unsigned long long>(p.preact) + static_cast<unsigned long long>(my_row * p.ldpre * esz)
                                        : 0ull;
    const long long boff = static_cast<long long>(ti.e) * p.n;
#pragma unroll 1
    for (int g = 0; g < BN / 64; ++g) {
      const int tcol = half * (BN / 2) + g * 32;
      const int col0 = ti.nb * BN + tcol;
      if (col0 >= p.n) break;
      const int valid = min(32, p.n - col0);
      float f[32];
      tmem_ld32_f(t_row + tcol, has_acc, f);
      if (p.bias != nullptr) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 8 < valid) {
            float b[8];
            if (p.bias_fp32)
              load8(reinterpret_cast<const float*>(p.bias) + boff + col0 + c * 8, b);
            else
              load8(reinterpret_cast<const __nv_bfloat16*>(p.bias) + boff + col0 + c * 8, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[c * 8 + i] = (p.bias_after_round ? bf16_round(f[c * 8 + i]) : f[c * 8 + i]) + b[i];
          }
        }
      }
      if (p.act != CSMOE_ACT_NONE || p.preact != nullptr) {
        if (!p.c_fp32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = bf16_round(f[i]);
        }
        if (p.preact != nullptr) {
          if (p.c_fp32)
            staged_store32<true>(stg, lane, f, my_pre, static_cast<long long>(col0) * 4, valid);
          else
            staged_store32<false>(stg, lane, f, my_pre, static_cast<long long>(col0) * 2, valid);
        }
        act_apply_vec<32>(f, p.act, !p.c_fp32);
      }
      if (p.c_fp32)
        staged_store32<true>(stg, lane, f, my_c, static_cast<long long>(col0) * 4, valid);
      else
        staged_store32<false>(stg, lane, f, my_c, static_cast<long long>(col0) * 2, valid);
    }
  } else if (p.epi == kEpiGluFwd) {
    // TMEM columns [0, BN/2) hold the gate and [BN/2, BN) the up projection of BN/2 output columns.  z = (gate | up) is
    // stored for the backward pass, h = up * silu(gate) feeds the down projection (Phi3MLP; each intermediate rounded
    // to bf16 like the eager reference).
    if constexpr (BN >= 256) {
      constexpr int kGate = BN / 2;
      const unsigned long long my_z = reinterpret_cast<unsigned long long>(p.preact) + static_cast<unsigned long long>(my_row * p.ldpre * 2);
#pragma unroll 1
      for (int g = 0; g < kGate / 64; ++g) {
        const int tcol = half * (kGate / 2) + g * 32;
        const int col0 = ti.nb * kGate + tcol;
        if (col0 >= p.glu_f) break;
        const int valid = min(32, p.glu_f - col0);
        float zg[32], zu[32];
        tmem_ld32_f(t_row + tcol, has_acc, zg);
#pragma unroll
        for (int i = 0; i < 32; ++i) zg[i] = bf16_round(zg[i]);
        staged_store32<false>(stg, lane, zg, my_z, static_cast<long long>(col0) * 2, valid);
        tmem_ld32_f(t_row + kGate + tcol, has_acc, zu);
#pragma unroll
        for (int i = 0; i < 32; ++i) zu[i] = bf16_round(zu[i]);
        staged_store32<false>(stg, lane, zu, my_z, static_cast<long long>(p.glu_f + col0) * 2, valid);
#pragma unroll
        for (int i = 0; i < 32; ++i) zu[i] *= bf16_round(act_apply(zg[i], CSMOE_ACT_SILU, true));
        staged_store32<false>(stg, lane, zu, my_c, static_cast<long long>(col0) * 2, valid);
      }
    }
  } else {
    // Backward epilogues: the accumulator is dh; multiply by the activation derivative at the saved z (bf16).
    const bool glu = p.epi == kEpiGluBwd;
    const unsigned long long my_z =
        row_ok ? reinterpret_cast<unsigned long long>(p.aux) + static_cast<unsigned long long>(my_row * p.ldaux * 2) : 0ull;
#pragma unroll 1
    for (int g = 0; g < BN / 64; ++g) {
      const int tcol = half * (BN / 2) + g * 32;
      const int col0 = ti.nb * BN + tcol;
      if (col0 >= p.n) break;
      const int valid = min(32, p.n - col0);
      float z0[32], dh[32];
      staged_load32_bf16(stg, lane, z0, my_z, static_cast<long long>(col0) * 2, valid);
      tmem_ld32_f(t_row + tcol, has_acc, dh);
      if (glu) {
        float z1[32];
        staged_load32_bf16(stg, lane, z1, my_z, static_cast<long long>(p.glu_f + col0) * 2, valid);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d = bf16_round(dh[i]);
          const float sg = bf16_round(act_apply(z0[i], CSMOE_ACT_SILU, true));
          z1[i] = bf16_round(d * z1[i]) * act_grad(z0[i], CSMOE_ACT_SILU, true);   // d gate
          dh[i] = d * sg;                                                      // d up
        }
        staged_store32<false>(stg, lane, z1, my_c, static_cast<long long>(col0) * 2, valid);
        staged_store32<false>(stg, lane, dh, my_c, static_cast<long long>(p.glu_f + col0) * 2, valid);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) dh[i] = bf16_round(dh[i]);
        act_grad_vec<32>(dh, z0, p.act, true);
        staged_store32<false>(stg, lane, dh, my_c, static_cast<long long>(col0) * 2, valid);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ TMA-store epilogue
// Plain (bias / activation / saved pre-activation) and fused-GLU epilogues whose outputs are row-major matrices: every
// 32 x 32 group is packed into the warp's staging tile and leaves through one TMA store, and the tcgen05.ld of group
// g+1 is in flight while group g is converted and staged (TMEM reads, 64 B/clk per SM, are what bounds this path).
template <bool FP32>
__device__ __forceinline__ void tma_store_packed(uint32_t stg, int lane, const uint32_t (&w)[FP32 ? 32 : 16],
                                                 const CUtensorMap* map, int col, int row, int e, uint32_t& slot) {
  // w: this thread's row of the group, 32 fp32 words (FP32) or 16 packed bf16 pairs
  constexpr int P = FP32 ? 128 : 64;
  const uint32_t buf = FP32 ? stg : stg + (slot & 1u) * 2048u;
  ++slot;
  if (lane == 0) {
    if (FP32)
      ptx::bulk_wait_group_read<0>();
    else
      ptx::bulk_wait_group_read<1>();
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < P / 16; ++c) sts128(stg_addr<P>(buf, lane, c), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
  ptx::fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    ptx::tma_store_3d(map, buf, col, row, e);
    ptx::bulk_commit_group();
  }
}

template <int MODE, int BN>
__device__ __forceinline__ void epilogue_tile_tma(const KParams& p, const Tile& ti, uint32_t t_row, bool has_acc, int row0,
                                                  int half, int lane, uint32_t stg, const CUtensorMap* map_c,
                                                  const CUtensorMap* map_p, uint32_t& slot) {
  const int te = MODE == CSMOE_GEMM_REDUCE ? ti.e : 0;
  if (p.epi == kEpiPlain) {
    constexpr int G = BN / 64;                        // 32-column groups per warp
    const long long boff = static_cast<long long>(ti.e) * p.n;
    const bool want_round = p.act != CSMOE_ACT_NONE || p.preact != nullptr;
    uint32_t v[32];
    if (has_acc) ptx::tmem_ld_32x32b_x32(t_row + half * (BN / 2), v);
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      const int tcol = half * (BN / 2) + g * 32;
      const int col0 = ti.nb * BN + tcol;
      const bool more = has_acc && g + 1 < G;
      if (has_acc) ptx::tmem_ld_wait_dep(v);
      if (col0 >= p.n) {
        if (more) ptx::tmem_ld_32x32b_x32(t_row + tcol + 32, v);
        continue;
      }
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = has_acc ? __uint_as_float(v[i]) : 0.f;
      if (p.bias != nullptr) {
        const int valid = min(32, p.n - col0);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 8 < valid) {
            float b[8];
            if (p.bias_fp32)
              load8(reinterpret_cast<const float*>(p.bias) + boff + col0 + c * 8, b);
            else
              load8(reinterpret_cast<const __nv_bfloat16*>(p.bias) + boff + col0 + c * 8, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[c * 8 + i] = (p.bias_after_round ? bf16_round(f[c * 8 + i]) : f[c * 8 + i]) + b[i];
          }
        }
      }
      if (p.c_fp32) {
        uint32_t w[32];
        if (want_round) {
          if (p.preact != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(f[i]);
            tma_store_packed<true>(stg, lane, w, map_p, col0, row0, te, slot);
          }
          act_apply_vec<32>(f, p.act, false);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(f[i]);
        if (more) ptx::tmem_ld_32x32b_x32(t_row + tcol + 32, v);   // streams in while this group is staged and stored
        tma_store_packed<true>(stg, lane, w, map_c, col0, row0, te, slot);
      } else {
        uint32_t w[16];
        if (want_round) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            w[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
            const float2 t = unpack_bf16(w[i]);
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
          }
          if (p.preact != nullptr) tma_store_packed<false>(stg, lane, w, map_p, col0, row0, te, slot);
          act_apply_vec<32>(f, p.act, true);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
        if (more) ptx::tmem_ld_32x32b_x32(t_row + tcol + 32, v);
        tma_store_packed<false>(stg, lane, w, map_c, col0, row0, te, slot);
      }
    }
  } else if (p.epi == kEpiGluFwd) {
    // TMEM columns [0, BN/2) = gate, [BN/2, BN) = up of BN/2 output columns; z = (gate | up) saved for backward,
    // h = up * silu(gate), every intermediate rounded to bf16 like the eager reference (Phi3MLP).
    if constexpr (BN >= 256) {
      constexpr int kGate = BN / 2;
      constexpr int G = kGate / 64;
      uint32_t vg[32], vu[32];
      if (has_acc) {
        ptx::tmem_ld_32x32b_x32(t_row + half * (kGate / 2), vg);
        ptx::tmem_ld_32x32b_x32(t_row + kGate + half * (kGate / 2), vu);
      }
#pragma unroll 1
      for (int g = 0; g < G; ++g) {
        const int tcol = half * (kGate / 2) + g * 32;
        const int col0 = ti.nb * kGate + tcol;
        uint32_t wg[16], wu[16];
        if (has_acc) {
          ptx::tmem_ld_wait_dep(vg);
          ptx::tmem_ld_wait_dep(vu);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            wg[i] = pack_bf16(__uint_as_float(vg[2 * i]), __uint_as_float(vg[2 * i + 1]));
            wu[i] = pack_bf16(__uint_as_float(vu[2 * i]), __uint_as_float(vu[2 * i + 1]));
          }
          if (g + 1 < G) {                               // next group's accumulators stream in during the SiLU math
            ptx::tmem_ld_32x32b_x32(t_row + tcol + 32, vg);
            ptx::tmem_ld_32x32b_x32(t_row + kGate + tcol + 32, vu);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) wg[i] = wu[i] = 0u;
        }
        if (col0 < p.glu_f) {
          tma_store_packed<false>(stg, lane, wg, map_p, col0, row0, 0, slot);
          tma_store_packed<false>(stg, lane, wu, map_p, p.glu_f + col0, row0, 0, slot);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 zg = unpack_bf16(wg[i]);
            const float2 zu = unpack_bf16(wu[i]);
            wg[i] = pack_bf16(zu.x * bf16_round(act_apply(zg.x, CSMOE_ACT_SILU, true)),
                              zu.y * bf16_round(act_apply(zg.y, CSMOE_ACT_SILU, true)));
          }
          tma_store_packed<false>(stg, lane, wg, map_c, col0, row0, 0, slot);
        }
      }
    }
  }
}

// Backward epilogues (dz = dh * act'(z), GLU: d gate | d up) with the saved pre-activation z arriving through TMA loads
// into the warp's staging tile instead of per-thread global loads: z of group g+1 is in flight while group g is
// computed, and the results leave through TMA stores.  Staging per warp (8 KiB): [0, 4 K) z (gate | up), [4 K, 8 K) the
// two output slots.  `zbar` is this warp's mbarrier, `zphase` its parity.
template <int BN>
__device__ __forceinline__ void epilogue_tile_bwd_tma(const KParams& p, const Tile& ti, uint32_t t_row, bool has_acc, int row0,
                                                      int half, int lane, uint32_t stg, const CUtensorMap* map_c,
                                                      const CUtensorMap* map_z, uint32_t zbar, uint32_t& zphase,
                                                      uint32_t& slot) {
  constexpr int G = BN / 64;
  const bool glu = p.epi == kEpiGluBwd;
  const uint32_t zin = stg, outb = stg + 4096u;
  const int colbase = ti.nb * BN + half * (BN / 2);
  if (lane == 0 && colbase < p.n) {
    ptx::mbar_arrive_expect_tx(zbar, glu ? 4096u : 2048u);
    ptx::tma_load_3d(zin, map_z, zbar, colbase, row0, 0);
    if (glu) ptx::tma_load_3d(zin + 2048u, map_z, zbar, p.glu_f + colbase, row0, 0);
  }
#pragma unroll 1
  for (int g = 0; g < G; ++g) {
    const int tcol = half * (BN / 2) + g * 32;
    const int col0 = ti.nb * BN + tcol;
    if (col0 >= p.n) break;
    uint32_t v[32];
    if (has_acc) ptx::tmem_ld_32x32b_x32(t_row + tcol, v);
    ptx::mbar_wait(zbar, zphase);
    zphase ^= 1u;
    uint32_t zg[16], zu[16];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 a = lds128(stg_addr<64>(zin, lane, c));
      zg[4 * c] = a.x, zg[4 * c + 1] = a.y, zg[4 * c + 2] = a.z, zg[4 * c + 3] = a.w;
    }
    if (glu) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 a = lds128(stg_addr<64>(zin + 2048u, lane, c));
        zu[4 * c] = a.x, zu[4 * c + 1] = a.y, zu[4 * c + 2] = a.z, zu[4 * c + 3] = a.w;
      }
    }
    __syncwarp();                                    // every lane has read z: the tile may be refilled
    if (lane == 0 && g + 1 < G && col0 + 32 < p.n) {
      ptx::mbar_arrive_expect_tx(zbar, glu ? 4096u : 2048u);
      ptx::tma_load_3d(zin, map_z, zbar, col0 + 32, row0, 0);
      if (glu) ptx::tma_load_3d(zin + 2048u, map_z, zbar, p.glu_f + col0 + 32, row0, 0);
    }
    if (has_acc) {
      ptx::tmem_ld_wait_dep(v);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0u;
    }
    if (glu) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 z0 = unpack_bf16(zg[i]), z1 = unpack_bf16(zu[i]);
        const float da = bf16_round(__uint_as_float(v[2 * i])), db = bf16_round(__uint_as_float(v[2 * i + 1]));
        const float sa = bf16_round(act_apply(z0.x, CSMOE_ACT_SILU, true)), sb = bf16_round(act_apply(z0.y, CSMOE_ACT_SILU, true));
        zu[i] = pack_bf16(da * sa, db * sb);                                                     // d up
        zg[i] = pack_bf16(bf16_round(da * z1.x) * act_grad(z0.x, CSMOE_ACT_SILU, true),
                          bf16_round(db * z1.y) * act_grad(z0.y, CSMOE_ACT_SILU, true));          // d gate
      }
      tma_store_packed<false>(outb, lane, zg, map_c, col0, row0, 0, slot);
      tma_store_packed<false>(outb, lane, zu, map_c, p.glu_f + col0, row0, 0, slot);
    } else {
      float dh[32], z0[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 z = unpack_bf16(zg[i]);
        z0[2 * i] = z.x, z0[2 * i + 1] = z.y;
        dh[2 * i] = bf16_round(__uint_as_float(v[2 * i]));
        dh[2 * i + 1] = bf16_round(__uint_as_float(v[2 * i + 1]));
      }
      act_grad_vec<32>(dh, z0, p.act, true);
#pragma unroll
      for (int i = 0; i < 16; ++i) zg[i] = pack_bf16(dh[2 * i], dh[2 * i + 1]);
      tma_store_packed<false>(outb, lane, zg, map_c, col0, row0, 0, slot);
    }
  }
}

template <int MODE, bool B_MN, int BN>
__global__ void __launch_bounds__(kThreads, 1)
grouped_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_p, const KParams p) {
  constexpr bool kAMn = (MODE == CSMOE_GEMM_REDUCE);
  constexpr bool kBMn = kAMn || B_MN;
  constexpr int kBBytes = BN * kBK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kStages = (BN == 256) ? 4 : 6;
  constexpr uint32_t kTmemCols = kAccStages * BN;
  constexpr uint32_t kIdesc = ptx::make_idesc_bf16(kBM, BN, kAMn, kBMn);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + kStages * kStageBytes;   // 8 epilogue warps x 4 KiB staging tiles
  const uint32_t bar_base = stg_base + kEpiWarps * kStageTileBytes;
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
            int kx = kb * kBK, ex = ti.e, arow = ti.a_row;
            if (p.kcat) {   // k loop over (expert, k-block): sum of the experts' products
              ex = kb / p.num_kb;
              kx = (kb % p.num_kb) * kBK;
              arow += ex * p.a_expert_rows;
            }
            ptx::tma_load_2d(sa, &tma_a, fb, kx, arow);
            if (!B_MN) {
              if (BN == 256 && p.epi == kEpiGluFwd) {
                // gate rows [nb*128, +128) then up rows [F + nb*128, +128): one output tile holds both halves
                ptx::tma_load_3d(sb, &tma_b, fb, kx, ti.nb * 128, ex);
                ptx::tma_load_3d(sb + 128 * kBK * 2, &tma_b, fb, kx, p.glu_f + ti.nb * 128, ex);
              } else {
                ptx::tma_load_3d(sb, &tma_b, fb, kx, ti.nb * BN, ex);
              }
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                ptx::tma_load_3d(sb + j * kSubTileBytes, &tma_b, fb, ti.nb * BN + j * 64, kx, ex);
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
    uint32_t acc = 0, acc_phase = 0, tma_slot = 0;
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
      if (p.tma_epi)
        epilogue_tile_tma<MODE, BN>(p, ti, t_row, has_acc, ti.mb * kBM + quad * 32, half, lane,
                                    stg_base + (warp - 2) * kStageTileBytes, &tma_c, &tma_p, tma_slot);
      else if (p.direct_epi)
        epilogue_tile<MODE, BN>(p, ti, t_row, has_acc, out_row, half);
      else
        epilogue_tile_staged<MODE, BN>(p, ti, t_row, has_acc, out_row - lane, half, lane, stg_base + (warp - 2) * kStageTileBytes);
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
    if (p.tma_epi && lane == 0) ptx::bulk_wait_group<0>();   // the staging tiles must outlive their TMA stores
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
__device__ __forceinline__ Tile decode_tile_pair(const KParams& p, long long t, int rank, int num_n = -1) {
  Tile ti;
  const int num_m2 = p.num_m_pairs;
  if (num_n < 0) num_n = p.num_n_blocks;
  if (MODE == CSMOE_GEMM_ROWS) {
    if (p.raster_m == 2) {
      // Bands aligned to the experts: the m-blocks of expert e are swept as ceil(w_e / band_cap) equal bands, m fastest
      // inside a band, so an expert's weights are fetched from DRAM once per band of ITS rows.  Fixed bands of 8
      // straddle the expert boundaries (4 experts x ~8.1 blocks at the bench shape) and re-read a whole expert for one
      // stray m-block: 0.96 GB of DRAM reads for 0.45 GB of operands in the fc1 launch (profiles/r02B_gemm_ncu_full.md).
      const int q = static_cast<int>(t / num_n);
      const int e = __ldg(p.tile_expert + 2 * q);
      if (e < 0) {              // past the last routed row
        ti.e = -1;
        ti.mb = q;
        ti.nb = 0;
        ti.a_row = q * 256 + rank * kBM;
        ti.b_row = 0;
        ti.nkb = p.num_kb;
        ti.valid = false;
        return ti;
      }
      const int m0 = __ldg(p.pad_offsets + e) >> 8, w_e = (__ldg(p.pad_offsets + e + 1) >> 8) - m0;
      const int r = static_cast<int>(t - static_cast<long long>(m0) * num_n);
      const int nbands = (w_e + p.band_cap - 1) / p.band_cap;
      const int bs = (w_e + nbands - 1) / nbands;
      const int band_tiles = bs * num_n;
      const int b = r / band_tiles, rr = r % band_tiles;
      const int mb0 = b * bs;
      const int w = min(bs, w_e - mb0);
      ti.nb = rr / w;
      ti.mb = m0 + mb0 + rr % w;
    } else if (p.raster_m) {
      // bands of `band` m-blocks, m fastest inside a band: neighbouring clusters ask for the same B tile at the same time
      const long long band_tiles = static_cast<long long>(p.band) * num_n;
      const int b = static_cast<int>(t / band_tiles);
      const int r = static_cast<int>(t % band_tiles);
      const int mb0 = b * p.band;
      const int w = min(p.band, num_m2 - mb0);
      ti.nb = r / w;
      ti.mb = mb0 + r % w;
    } else {
      const long long band_tiles = static_cast<long long>(p.band) * num_m2;
      const int b = static_cast<int>(t / band_tiles);
      const int r = static_cast<int>(t % band_tiles);
      const int nb0 = b * p.band;
      const int w = min(p.band, num_n - nb0);
      ti.mb = r / w;
      ti.nb = nb0 + r % w;
    }
    if (p.kcat) {
      ti.e = 0;
      ti.a_row = ti.mb * 256 + rank * kBM;
    } else if (p.dense) {
      const int dm2 = p.dense_mblocks / 2;
      ti.e = ti.mb / dm2;
      ti.a_row = (ti.mb % dm2) * 256 + rank * kBM + ti.e * p.a_expert_rows;
    } else {
      ti.e = __ldg(p.tile_expert + 2 * ti.mb);
      ti.a_row = ti.mb * 256 + rank * kBM;
    }
    ti.b_row = 0;
    ti.nkb = p.kcat ? p.num_kb * p.num_experts : p.num_kb;
    ti.valid = ti.e >= 0;
  } else {
    const long long per_e = static_cast<long long>(num_m2) * num_n;
    ti.e = static_cast<int>(t / per_e);
    const int r = static_cast<int>(t % per_e);
    const int band_tiles = p.band * num_m2;
    const int b = r / band_tiles, rr = r % band_tiles;
    const int nb0 = b * p.band;
    const int w = min(p.band, num_n - nb0);
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

// CL = CTAs per cluster.  CL == 2: one pair.  CL == 4: two pairs that work on the same 256 rows and on neighbouring
// n-blocks (nb = 2 * nb2 + pair); every CTA fetches only half of its 128 x 64 A sub-tile and multicasts it to the CTA at
// the same position in the other pair, so the L2 -> SM traffic per CTA and k-block drops from 32 to 24 KiB (the pair
// kernel's limiter) while both pairs keep their double-buffered accumulators.  A stage is then
// written by loads of both pairs, so the empty barriers count one tcgen05.commit per pair (multicast to all four CTAs)
// and the two pairs advance through the k loop in lockstep.
template <int MODE, bool B_MN, int CL>
__device__ __forceinline__ void pair_kernel_body(const CUtensorMap& tma_a, const CUtensorMap& tma_b, const CUtensorMap& tma_c,
                                                 const CUtensorMap& tma_p, const CUtensorMap& tma_z, const KParams& p) {
  static_assert(CL == 2 || CL == 4, "cluster of one or two CTA pairs");
  constexpr int BN = 256;
  constexpr bool kAMn = (MODE == CSMOE_GEMM_REDUCE);
  constexpr bool kBMn = kAMn || B_MN;
  constexpr uint32_t kTmemCols = kAccStages * BN;
  constexpr uint32_t kIdesc = ptx::make_idesc_bf16(256, BN, kAMn, kBMn);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int n_stages = p.stages;
  const uint32_t stg_base = smem_base + n_stages * kPairStageBytes;
  const uint32_t bar_base = stg_base + kEpiWarps * p.stg_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kPairStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kPairStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kPairStages + kAccStages + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kPairStages + 2 * kAccStages);
  auto z_bar = [&](int w) { return bar_base + 8u * (2 * kPairStages + 2 * kAccStages + 2 + w); };
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = static_cast<int>(ptx::cluster_ctarank());   // 0 .. CL-1
  const int rank = crank & 1;                                    // position inside the pair
  const int pair = crank >> 1;                                   // which pair of the cluster
  const bool leader = rank == 0;
  const long long cluster_id = blockIdx.x / CL;
  const long long num_clusters = gridDim.x / CL;
  const int num_n = CL == 4 ? (p.num_n_blocks + 1) / 2 : p.num_n_blocks;     // n-block groups a cluster iterates over
  const uint16_t pair_mask = static_cast<uint16_t>(0x3u << (2 * pair));      // the two CTAs of this pair
  const uint16_t all_mask = static_cast<uint16_t>((1u << CL) - 1u);
  auto tile_of = [&](long long t) {
    Tile ti = decode_tile_pair<MODE>(p, t, rank, num_n);
    if (CL == 4) ti.nb = 2 * ti.nb + pair;
    return ti;
  };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kPairStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), CL / 2);
    }
    for (int a = 0; a < kAccStages; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), 2 * kEpiWarps);
    }
    for (int w = 0; w < kEpiWarps; ++w) ptx::mbar_init(z_bar(w), 1);
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
  const bool st_on = p.stats != nullptr;
  const long long st_t0 = st_on ? clock64() : 0;
  const unsigned long long st_g0 = st_on ? ptx::globaltimer() : 0ull;

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      unsigned long long w_empty = 0;
      int dbg_filled = 0;
      for (long long t = cluster_id; t < p.total_tiles; t += num_clusters) {
        const Tile ti = tile_of(t);
        if (!ti.valid) continue;
        for (int kb = 0; kb < ti.nkb; ++kb) {
          timed_wait(empty_bar(stage), phase ^ 1u, st_on, w_empty);
          if ((p.dbg_mode & 1) && dbg_filled >= n_stages) {
            if (leader) ptx::mbar_arrive(full_bar(stage));
            if (++stage == n_stages) {
              stage = 0;
              phase ^= 1u;
            }
            continue;
          }
          ++dbg_filled;
          if (leader) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * kPairStageBytes);
          // the pair leader's barrier collects both CTAs' bytes
          const uint32_t fb = CL == 2 ? ptx::mapa(full_bar(stage), 0) : (full_bar(stage) & ptx::kPeerBitMask);
          const uint32_t sa = smem_base + stage * kPairStageBytes;
          const uint32_t sb = sa + kABytes;
          if (MODE == CSMOE_GEMM_ROWS) {
            int kx = kb * kBK, ex = ti.e, arow = ti.a_row;
            if (p.kcat) {
              ex = kb / p.num_kb;
              kx = (kb % p.num_kb) * kBK;
              arow += ex * p.a_expert_rows;
            }
            if (CL == 2) {
              ptx::tma_load_2d_cg2(sa, &tma_a, fb, kx, arow);
            } else {   // this CTA's half of the 128 rows, delivered to both pairs
              ptx::tma_load_2d_cg2_mc(sa + pair * (kABytes / 2), &tma_a, fb, kx, arow + pair * 64,
                                      static_cast<uint16_t>((1u << crank) | (1u << (crank ^ 2))));
            }
            if (!B_MN) {
              const int brow = p.epi == kEpiGluFwd ? (rank == 0 ? ti.nb * 128 : p.glu_f + ti.nb * 128)
                                                   : ti.nb * BN + rank * 128;
              ptx::tma_load_3d_cg2(sb, &tma_b, fb, kx, brow, ex);
            } else {
#pragma unroll
              for (int j = 0; j < 2; ++j)
                ptx::tma_load_3d_cg2(sb + j * kSubTileBytes, &tma_b, fb, ti.nb * BN + rank * 128 + j * 64, kx, ex);
            }
          } else {
            if (CL == 2) {
#pragma unroll
              for (int j = 0; j < 2; ++j)
                ptx::tma_load_2d_cg2(sa + j * kSubTileBytes, &tma_a, fb, ti.mb * 256 + rank * 128 + j * 64,
                                     ti.a_row + kb * kBK);
            } else {
              ptx::tma_load_2d_cg2_mc(sa + pair * kSubTileBytes, &tma_a, fb, ti.mb * 256 + rank * 128 + pair * 64,
                                      ti.a_row + kb * kBK, static_cast<uint16_t>((1u << crank) | (1u << (crank ^ 2))));
            }
#pragma unroll
            for (int j = 0; j < 2; ++j)
              ptx::tma_load_2d_cg2(sb + j * kSubTileBytes, &tma_b, fb, ti.nb * BN + rank * 128 + j * 64,
                                   ti.b_row + kb * kBK);
          }
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (st_on) p.stats[blockIdx.x * 8 + 0] = w_empty;
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      unsigned long long w_full = 0, w_tempty = 0, n_tiles = 0;
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (long long t = cluster_id; t < p.total_tiles; t += num_clusters) {
        const Tile ti = tile_of(t);
        if (!ti.valid || ti.nkb == 0) continue;
        ++n_tiles;
        timed_wait(tempty_bar(acc), acc_phase ^ 1u, st_on, w_tempty);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < ti.nkb; ++kb) {
          timed_wait(full_bar(stage), phase, st_on, w_full);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * kPairStageBytes;
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t adesc = kAMn ? ptx::make_smem_desc_sw128(sa + k * 2048, kSubTileBytes, 1024)
                                        : ptx::make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t bdesc = kBMn ? ptx::make_smem_desc_sw128(sb + k * 2048, kSubTileBytes, 1024)
                                        : ptx::make_smem_desc_sw128(sb + k * 32, 16, 1024);
            if (!(p.dbg_mode & 2)) ptx::umma_f16_cg2(d_tmem, adesc, bdesc, kIdesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit_cg2_mc(empty_bar(stage), all_mask);
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        ptx::umma_commit_cg2_mc(tfull_bar(acc), pair_mask);
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
      if (st_on) {
        p.stats[blockIdx.x * 8 + 1] = w_full;
        p.stats[blockIdx.x * 8 + 2] = w_tempty;
        p.stats[blockIdx.x * 8 + 5] = n_tiles;
      }
    }
  } else {
    // ===================================================== epilogue (both CTAs; 8 warps each)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quad * 32 + lane;
    unsigned long long epi_cycles = 0;
    uint32_t tma_slot = 0, z_phase = 0;
    uint32_t acc = 0, acc_phase = 0;
    for (long long t = cluster_id; t < p.total_tiles; t += num_clusters) {
      const Tile ti = tile_of(t);
      if (!ti.valid) continue;
      const bool has_acc = ti.nkb > 0;
      if (has_acc) {
        ptx::mbar_wait(tfull_bar(acc), acc_phase);
        ptx::tc_fence_after();
      }
      const long long out_row = static_cast<long long>(ti.mb) * 256 + rank * kBM + row_in_tile;
      const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
      const long long e0 = st_on ? clock64() : 0;
      if (MODE == CSMOE_GEMM_ROWS && p.tma_epi == 2)
        epilogue_tile_bwd_tma<BN>(p, ti, t_row, has_acc, ti.mb * 256 + rank * kBM + quad * 32, half, lane,
                                  stg_base + (warp - 2) * p.stg_bytes, &tma_c, &tma_z, z_bar(warp - 2), z_phase, tma_slot);
      else if (p.tma_epi)
        epilogue_tile_tma<MODE, BN>(p, ti, t_row, has_acc, ti.mb * 256 + rank * kBM + quad * 32, half, lane,
                                    stg_base + (warp - 2) * p.stg_bytes, &tma_c, &tma_p, tma_slot);
      else if (p.direct_epi)
        epilogue_tile<MODE, BN>(p, ti, t_row, has_acc, out_row, half);
      else
        epilogue_tile_staged<MODE, BN>(p, ti, t_row, has_acc, out_row - lane, half, lane, stg_base + (warp - 2) * p.stg_bytes);
      if (has_acc) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader)
            ptx::mbar_arrive(tempty_bar(acc));
          else
            ptx::mbar_arrive_remote_relaxed(ptx::mapa(tempty_bar(acc), crank & ~1));
        }
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
      if (st_on) epi_cycles += static_cast<unsigned long long>(clock64() - e0);
    }
    if (p.tma_epi && lane == 0) ptx::bulk_wait_group<0>();   // the staging tiles must outlive their TMA stores
    if (st_on && warp == 2 && lane == 0) p.stats[blockIdx.x * 8 + 3] = epi_cycles;
  }

  if (st_on && threadIdx.x == 0) {
    p.stats[blockIdx.x * 8 + 4] = static_cast<unsigned long long>(clock64() - st_t0);
    p.stats[blockIdx.x * 8 + 6] = ptx::globaltimer() - st_g0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();  // the peer's smem / barriers must outlive the leader's last MMA and commit
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
}

template <int MODE, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
grouped_gemm_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                         const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_p,
                         const __grid_constant__ CUtensorMap tma_z, const KParams p) {
  pair_kernel_body<MODE, B_MN, 2>(tma_a, tma_b, tma_c, tma_p, tma_z, p);
}

template <int MODE, bool B_MN>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(kThreads, 1)
grouped_gemm_quad_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                         const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_p,
                         const __grid_constant__ CUtensorMap tma_z, const KParams p) {
  pair_kernel_body<MODE, B_MN, 4>(tma_a, tma_b, tma_c, tma_p, tma_z, p);
}


// ------------------------------------------------------------------------------------------------ wide CTA-pair kernel
// The 256 x 256 pair kernel moves 64 KiB from L2 to the two SMs per 64-deep k-block: at cuBLAS-level tensor rates that
// is 11-12 TB/s, the practical ceiling of the L2 -> SM fabric on this part (ncu: lts2xbar 10.5 TB/s, the kernel's
// limiter; profiles/r01c_gemm_l2_bound.md).  This variant gives each pair a 256 x 512 tile: every A k-block is reused
// for two 256-column MMAs, so the traffic per flop drops by a quarter (48 KiB per CTA per k-block for twice the math).
// The accumulator takes all 512 TMEM columns, so the epilogue of a tile is no longer hidden behind the next tile's
// MMAs (TMA prefetch of the next tile still overlaps it); that costs ~4k cycles per tile (TMEM reads at 64 B/clk)
// against >= 32k cycles of MMAs.
constexpr int kWideStages = 4;
constexpr int kWideHalfBytes = 128 * kBK * 2;                       // one CTA's share of one 256-column half: 16 KiB
constexpr int kWideStageBytes = kABytes + 2 * kWideHalfBytes;       // 48 KiB per CTA per stage

template <int MODE, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
grouped_gemm_wide_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                         const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_p,
                         const KParams p) {
  constexpr int BN = 512;
  constexpr bool kAMn = (MODE == CSMOE_GEMM_REDUCE);
  constexpr bool kBMn = kAMn || B_MN;
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kIdesc = ptx::make_idesc_bf16(256, 256, kAMn, kBMn);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + kWideStages * kWideStageBytes;
  const uint32_t bar_base = stg_base + kEpiWarps * kStageTileBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWideStages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * kWideStages);
  const uint32_t tempty_bar = bar_base + 8u * (2 * kWideStages + 1);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kWideStages + 2);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(ptx::cluster_ctarank());
  const bool leader = rank == 0;
  const long long cluster_id = blockIdx.x >> 1;
  const long long num_clusters = gridDim.x >> 1;
  const bool glu = p.epi == kEpiGluFwd;
  // halves of the tile that hold valid output columns (the last n-block of a plain GEMM may need only one)
  auto tile_halves = [&](const Tile& ti) { return (glu || ti.nb * BN + 256 < p.n) ? 2 : 1; };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWideStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(tfull_bar, 1);
    ptx::mbar_init(tempty_bar, 2 * kEpiWarps);
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
  const bool st_on = p.stats != nullptr;
  const long long st_t0 = st_on ? clock64() : 0;
  const unsigned long long st_g0 = st_on ? ptx::globaltimer() : 0ull;

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      unsigned long long w_empty = 0;
      int dbg_filled = 0;
      for (long long t = cluster_id; t < p.total_tiles; t += num_clusters) {
        const Tile ti = decode_tile_pair<MODE>(p, t, rank);
        if (!ti.valid) continue;
        const int halves = tile_halves(ti);
        const uint32_t bytes = 2u * (kABytes + halves * kWideHalfBytes);
        for (int kb = 0; kb < ti.nkb; ++kb) {
          timed_wait(empty_bar(stage), phase ^ 1u, st_on, w_empty);
          if ((p.dbg_mode & 1) && dbg_filled >= kWideStages) {
            if (leader) ptx::mbar_arrive(full_bar(stage));
            if (++stage == kWideStages) {
              stage = 0;
              phase ^= 1u;
            }
            continue;
          }
          ++dbg_filled;
          if (leader) ptx::mbar_arrive_expect_tx(full_bar(stage), bytes);
          const uint32_t fb = ptx::mapa(full_bar(stage), 0);
          const uint32_t sa = smem_base + stage * kWideStageBytes;
          const uint32_t sb = sa + kABytes;
          if (MODE == CSMOE_GEMM_ROWS) {
            int kx = kb * kBK, ex = ti.e, arow = ti.a_row;
            if (p.kcat) {
              ex = kb / p.num_kb;
              kx = (kb % p.num_kb) * kBK;
              arow += ex * p.a_expert_rows;
            }
            ptx::tma_load_2d_cg2(sa, &tma_a, fb, kx, arow);
            for (int h = 0; h < halves; ++h) {
              if (!B_MN) {
                const int brow = glu ? h * p.glu_f + ti.nb * 256 + rank * 128 : ti.nb * BN + h * 256 + rank * 128;
                ptx::tma_load_3d_cg2(sb + h * kWideHalfBytes, &tma_b, fb, kx, brow, ex);
              } else {
#pragma unroll
                for (int j = 0; j < 2; ++j)
                  ptx::tma_load_3d_cg2(sb + h * kWideHalfBytes + j * kSubTileBytes, &tma_b, fb,
                                       ti.nb * BN + h * 256 + rank * 128 + j * 64, kx, ex);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              ptx::tma_load_2d_cg2(sa + j * kSubTileBytes, &tma_a, fb, ti.mb * 256 + rank * 128 + j * 64,
                                   ti.a_row + kb * kBK);
            for (int h = 0; h < halves; ++h)
#pragma unroll
              for (int j = 0; j < 2; ++j)
                ptx::tma_load_2d_cg2(sb + h * kWideHalfBytes + j * kSubTileBytes, &tma_b, fb,
                                     ti.nb * BN + h * 256 + rank * 128 + j * 64, ti.b_row + kb * kBK);
          }
          if (++stage == kWideStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (st_on) p.stats[blockIdx.x * 8 + 0] = w_empty;
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      unsigned long long w_full = 0, w_tempty = 0, n_tiles = 0;
      uint32_t stage = 0, phase = 0, acc_phase = 0;
      for (long long t = cluster_id; t < p.total_tiles; t += num_clusters) {
        const Tile ti = decode_tile_pair<MODE>(p, t, rank);
        if (!ti.valid || ti.nkb == 0) continue;
        ++n_tiles;
        const int halves = tile_halves(ti);
        timed_wait(tempty_bar, acc_phase ^ 1u, st_on, w_tempty);
        ptx::tc_fence_after();
        for (int kb = 0; kb < ti.nkb; ++kb) {
          timed_wait(full_bar(stage), phase, st_on, w_full);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * kWideStageBytes;
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t adesc = kAMn ? ptx::make_smem_desc_sw128(sa + k * 2048, kSubTileBytes, 1024)
                                        : ptx::make_smem_desc_sw128(sa + k * 32, 16, 1024);
            for (int h = 0; h < halves; ++h) {
              const uint32_t sbh = sb + h * kWideHalfBytes;
              const uint64_t bdesc = kBMn ? ptx::make_smem_desc_sw128(sbh + k * 2048, kSubTileBytes, 1024)
                                          : ptx::make_smem_desc_sw128(sbh + k * 32, 16, 1024);
              if (!(p.dbg_mode & 2)) ptx::umma_f16_cg2(tmem_base + h * 256, adesc, bdesc, kIdesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          ptx::umma_commit_cg2_mc(empty_bar(stage), 0x3);
          if (++stage == kWideStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        ptx::umma_commit_cg2_mc(tfull_bar, 0x3);
        acc_phase ^= 1u;
      }
      if (st_on) {
        p.stats[blockIdx.x * 8 + 1] = w_full;
        p.stats[blockIdx.x * 8 + 2] = w_tempty;
        p.stats[blockIdx.x * 8 + 5] = n_tiles;
      }
    }
  } else {
    // ===================================================== epilogue (both CTAs; 8 warps each, 256 columns per warp)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quad * 32 + lane;
    unsigned long long epi_cycles = 0;
    uint32_t tma_slot = 0;
    uint32_t acc_phase = 0;
    for (long long t = cluster_id; t < p.total_tiles; t += num_clusters) {
      const Tile ti = decode_tile_pair<MODE>(p, t, rank);
      if (!ti.valid) continue;
      const bool has_acc = ti.nkb > 0;
      if (has_acc) {
        ptx::mbar_wait(tfull_bar, acc_phase);
        ptx::tc_fence_after();
      }
      const long long out_row = static_cast<long long>(ti.mb) * 256 + rank * kBM + row_in_tile;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      const long long e0 = st_on ? clock64() : 0;
      if (p.tma_epi)
        epilogue_tile_tma<MODE, BN>(p, ti, t_row, has_acc, ti.mb * 256 + rank * kBM + quad * 32, half, lane,
                                    stg_base + (warp - 2) * kStageTileBytes, &tma_c, &tma_p, tma_slot);
      else if (p.direct_epi)
        epilogue_tile<MODE, BN>(p, ti, t_row, has_acc, out_row, half);
      else
        epilogue_tile_staged<MODE, BN>(p, ti, t_row, has_acc, out_row - lane, half, lane, stg_base + (warp - 2) * kStageTileBytes);
      if (has_acc) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader)
            ptx::mbar_arrive(tempty_bar);
          else
            ptx::mbar_arrive_remote_relaxed(ptx::mapa(tempty_bar, 0));
        }
        acc_phase ^= 1u;
      }
      if (st_on) epi_cycles += static_cast<unsigned long long>(clock64() - e0);
    }
    if (p.tma_epi && lane == 0) ptx::bulk_wait_group<0>();   // the staging tiles must outlive their TMA stores
    if (st_on && warp == 2 && lane == 0) p.stats[blockIdx.x * 8 + 3] = epi_cycles;
  }

  if (st_on && threadIdx.x == 0) {
    p.stats[blockIdx.x * 8 + 4] = static_cast<unsigned long long>(clock64() - st_t0);
    p.stats[blockIdx.x * 8 + 6] = ptx::globaltimer() - st_g0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
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

// Output map for the TMA-store epilogue: [experts][rows][cols] of bf16 / fp32, 32 x 32 boxes, swizzle = the staging
// tile's (64-byte rows for bf16, 128-byte rows for fp32).
int encode_out_map(CUtensorMap* map, const void* base, bool fp32, long long cols, long long rows, long long experts,
                   long long ld, long long expert_stride) {
  EncodeFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return CSMOE_ERR_DRIVER;
  }
  const cuuint64_t esz = fp32 ? 4 : 2;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)experts};
  cuuint64_t str[2] = {(cuuint64_t)ld * esz, (cuuint64_t)(experts > 1 ? expert_stride : rows * ld) * esz};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base),
                  dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  fp32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (output map) failed: CUresult %d (dims %lld,%lld,%lld ld %lld)", (int)r, cols, rows,
              experts, ld);
    return CSMOE_ERR_DRIVER;
  }
  return CSMOE_OK;
}

struct Maps {
  CUtensorMap a, b, c, p, z;
  CUtensorMap aq;   // A with half-height boxes (quad clusters); swapped into `a` when that kernel is chosen
};

template <int MODE, bool B_MN, int BN>
int launch(const Maps& m, const KParams& kp, int grid, cudaStream_t stream) {
  constexpr int kStages = (BN == 256) ? 4 : 6;
  constexpr int kSmem = kStages * (kABytes + BN * kBK * 2) + kEpiWarps * kStageTileBytes + 1024 + 256;
  auto kern = grouped_gemm_kernel<MODE, B_MN, BN>;
  static bool configured = false;  // per instantiation; benign race (idempotent attribute set)
  if (!configured) {
    CSMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  kern<<<grid, kThreads, kSmem, stream>>>(m.a, m.b, m.c, m.p, kp);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

template <int MODE, bool B_MN>
int launch_pair(const Maps& m, const KParams& kp, int clusters, cudaStream_t stream) {
  // 6 stages + 4 KiB of staging per epilogue warp, or 5 stages + 8 KiB (TMA-fed backward epilogue)
  constexpr int kSmem = (kPairStages - 1) * kPairStageBytes + kEpiWarps * 2 * kStageTileBytes + 1024 + 256;
  static_assert(kSmem >= kPairStages * kPairStageBytes + kEpiWarps * kStageTileBytes + 1024 + 256, "pair kernel smem");
  auto kern = grouped_gemm_pair_kernel<MODE, B_MN>;
  static bool configured = false;
  if (!configured) {
    CSMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  kern<<<2 * clusters, kThreads, kSmem, stream>>>(m.a, m.b, m.c, m.p, m.z, kp);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

// Clusters of 4 (two pairs sharing A through multicast).  Returns the number of co-resident clusters through `cap` when
// `clusters` <= 0 (query only).
template <int MODE, bool B_MN>
int launch_quad(const Maps& m, const KParams& kp, int clusters, cudaStream_t stream, int* cap = nullptr) {
  constexpr int kSmem = (kPairStages - 1) * kPairStageBytes + kEpiWarps * 2 * kStageTileBytes + 1024 + 256;
  auto kern = grouped_gemm_quad_kernel<MODE, B_MN>;
  static bool configured = false;
  static int max_clusters = 0;
  if (!configured) {
    CSMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(4 * 64);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    CSMOE_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    max_clusters = n;
    configured = true;
  }
  if (cap != nullptr) *cap = max_clusters;
  if (clusters <= 0) return CSMOE_OK;
  kern<<<4 * clusters, kThreads, kSmem, stream>>>(m.a, m.b, m.c, m.p, m.z, kp);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

template <int MODE, bool B_MN>
int launch_wide(const Maps& m, const KParams& kp, int clusters, cudaStream_t stream) {
  constexpr int kSmem = kWideStages * kWideStageBytes + kEpiWarps * kStageTileBytes + 1024 + 256;
  auto kern = grouped_gemm_wide_kernel<MODE, B_MN>;
  static bool configured = false;
  if (!configured) {
    CSMOE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  kern<<<2 * clusters, kThreads, kSmem, stream>>>(m.a, m.b, m.c, m.p, kp);
  CSMOE_CHECK_LAUNCH();
  return CSMOE_OK;
}

// CSMOE_GEMM_WIDE: bit 0 = ROWS launches, bit 1 = REDUCE launches may use the 256 x 512 kernel (default 3; 0 = never)
int wide_mask() {
  static const int m = []() {
    const char* v = getenv("CSMOE_GEMM_WIDE");
    return v == nullptr ? 3 : atoi(v);
  }();
  return m;
}

// CSMOE_GEMM_STATS=1: per-role wait-cycle counters of the pair / wide kernels, printed to stderr after every launch
// (synchronises the stream: a bring-up / tuning aid, never on in production).
unsigned long long* stats_buffer() {
  static unsigned long long* buf = []() -> unsigned long long* {
    const char* v = getenv("CSMOE_GEMM_STATS");
    if (v == nullptr || v[0] == '0') return nullptr;
    void* p = nullptr;
    if (cudaMalloc(&p, 256 * 8 * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
    return static_cast<unsigned long long*>(p);
  }();
  return buf;
}

void print_stats(const char* what, const KParams& kp, int ctas, cudaStream_t stream) {
  unsigned long long h[256 * 8];
  cudaStreamSynchronize(stream);
  cudaMemcpy(h, kp.stats, sizeof(unsigned long long) * 8 * ctas, cudaMemcpyDeviceToHost);
  double s[6] = {0, 0, 0, 0, 0, 0}, ep = 0, tot_max = 0;
  for (int c = 0; c < ctas; ++c) {
    if (c % 2 == 0) {
      for (int i = 0; i < 6; ++i) s[i] += static_cast<double>(h[c * 8 + i]);
    } else {
      ep += static_cast<double>(h[c * 8 + 3]);
    }
    if (static_cast<double>(h[c * 8 + 4]) > tot_max) tot_max = static_cast<double>(h[c * 8 + 4]);
  }
  const double n = ctas / 2;
  fprintf(stderr, "[csmoe gemm stats] SM clock %.0f MHz; ", 1e3 * static_cast<double>(h[4]) / static_cast<double>(h[6] ? h[6] : 1));
  fprintf(stderr,
          "%s tiles/cluster %.2f total %.0f cyc (max %.0f) | leader: producer-wait-empty %.1f%% "
          "mma-wait-full %.1f%% mma-wait-tempty %.1f%% epilogue-busy %.1f%% (peer %.1f%%) | epilogue cyc/tile %.0f\n",
          what, s[5] / n, s[4] / n, tot_max, 100 * s[0] / s[4], 100 * s[1] / s[4], 100 * s[2] / s[4], 100 * s[3] / s[4],
          100 * ep / s[4], s[3] / (s[5] > 0 ? s[5] : 1));
}

// Epilogue store path.  Measured on B200 (scripts/gemm_bench.py, profiles/r01d_gemm_tuning.md): the staged, coalesced
// stores win where a tile writes several outputs (fused GLU forward: z gate, z up and h, +7 %), are neutral for the
// plain single-output epilogue and lose ~9 % in wgrad, whose main loop is shared-memory-bandwidth bound and feels the
// extra st.shared / ld.shared traffic.  CSMOE_GEMM_EPI=direct|staged forces one path for A/B runs.
int epilogue_override() {
  static const int v = []() {
    const char* e = getenv("CSMOE_GEMM_EPI");
    if (e == nullptr) return 0;
    return e[0] == 'd' ? 1 : (e[0] == 's' ? 2 : (e[0] == 't' ? 3 : 0));
  }();
  return v;
}

// CSMOE_GEMM_TMA: bit 0 = ROWS, bit 1 = REDUCE launches use the TMA-store epilogue by default (default 3)
bool tma_default(int mode) {
  static const int m = []() {
    const char* v = getenv("CSMOE_GEMM_TMA");
    return v == nullptr ? 3 : atoi(v);
  }();
  return (m & (mode == CSMOE_GEMM_ROWS ? 1 : 2)) != 0;
}

// CSMOE_GEMM_QUAD: bit 0 = ROWS, bit 1 = REDUCE launches of the 256 x 256 kernel run as clusters of two pairs with the A
// operand multicast (default 0 until measured; see pair_kernel_body)
int quad_mask() {
  static const int m = []() {
    const char* v = getenv("CSMOE_GEMM_QUAD");
    return v == nullptr ? 0 : atoi(v);
  }();
  return m;
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
  if (a->rowsum != nullptr) {
    CSMOE_CHECK_ARG(a->mode == CSMOE_GEMM_ROWS && !glu_fwd && !act_bwd && a->c_rows == nullptr && a->c != nullptr,
                    "csmoe_grouped_gemm: rowsum needs a ROWS launch with the plain epilogue and a local C");
  }
  if (a->sum_experts) {
    CSMOE_CHECK_ARG(a->mode == CSMOE_GEMM_ROWS && a->dense && a->a_expert_rows > 0 && a->bias == nullptr && !glu_fwd &&
                        !act_bwd && a->preact == nullptr && a->c_rows == nullptr && a->k % kBK == 0,
                    "csmoe_grouped_gemm: sum_experts needs a dense ROWS launch with per-expert A rows, the plain epilogue "
                    "and k a multiple of 64");
  }

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

  // clusters of two pairs (A multicast): plain pair launches only (no per-row destinations; the n-blocks come in twos)
  const bool quad = CSMOE_BUILD_QUAD != 0 && pair && (quad_mask() & (a->mode == CSMOE_GEMM_ROWS ? 1 : 2)) != 0 && a->c_rows == nullptr;

  KParams kp{};
  kp.n = static_cast<int>(a->n);
  kp.num_experts = E;
  kp.dense = a->dense;
  kp.kcat = a->sum_experts ? 1 : 0;
  kp.rowsum = a->rowsum;
  kp.rowsum_ld = static_cast<int>((a->n + 63) / 64);
  kp.rowsum_round = a->rowsum_round;
  kp.dense_mblocks = a->dense ? static_cast<int>(a->dense_rows / kBM) : 0;
  kp.dense_kblocks = a->dense ? static_cast<int>(a->dense_rows / kBK) : 0;
  kp.a_expert_rows = static_cast<int>(a->a_expert_rows);
  kp.act = a->act;
  kp.c_fp32 = a->c_dtype == CSMOE_F32;
  kp.bias_fp32 = a->bias_dtype == CSMOE_F32;
  kp.bias_after_round = a->bias_after_round != 0;
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
  {
    static const int band = []() { const char* v = getenv("CSMOE_GEMM_BAND"); return v ? atoi(v) : 8; }();
    kp.band = band > 0 ? band : 8;
    // m-fastest raster (neighbouring clusters share the B tile, one expert's weights stay hot in L2 while its row
    // band is swept): DRAM reads of the fc1 launch 1.45 -> 0.75 GB, +7 % (profiles/r01k_gemm_epilogue.md).
    static const int rm = []() { const char* v = getenv("CSMOE_GEMM_RASTER"); return v ? atoi(v) : 2; }();
    // ... only where the expert changes along m: with one local expert (expert-parallel ranks of the bench shape) every
    // tile shares the same weights and the n-fastest bands are 5 % faster (EP4 step 3.88 vs 3.70 ms).
    kp.raster_m = E >= 2 ? rm : 0;
    // expert-aligned bands (CSMOE_GEMM_RASTER=2, the default): routed launches whose segments are 256-row aligned
    static const int cap = []() { const char* v = getenv("CSMOE_GEMM_BAND_CAP"); return v ? atoi(v) : 12; }();
    kp.band_cap = cap > 0 ? cap : 12;
    if (kp.raster_m == 2 && !(a->mode == CSMOE_GEMM_ROWS && !a->dense && a->pad_offsets != nullptr && a->tile_expert != nullptr &&
                              a->row_tile >= 256))
      kp.raster_m = 1;
  }
  kp.aux = a->aux;
  kp.ldaux = a->ldaux;
  kp.c_rows = reinterpret_cast<const unsigned long long*>(a->c_rows);
  kp.direct_epi = 1;  // decided once the epilogue kind is known (below)
  {
    static const int dbg = []() { const char* v = getenv("CSMOE_GEMM_DBG"); return v ? atoi(v) : 0; }();
    kp.dbg_mode = dbg;
  }
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
  // ... and where rows leave the GPU (expert-parallel return, c_rows): 64-byte row segments instead of 16-byte ones
  // make far better NVLink packets.
  // ... and for every ROWS launch with an activation / second output or a short k loop (k < 2048: epilogue-heavy;
  // SigLIP fc1, k = 1152: 0.32 vs 0.41 ms with GELU + saved pre-activation, 0.225 vs 0.242 ms plain).
  kp.direct_epi = (kp.epi == kEpiGluFwd || a->c_rows != nullptr ||
                   (a->mode == CSMOE_GEMM_ROWS && (a->preact != nullptr || a->act != CSMOE_ACT_NONE || act_bwd || a->k < 2048)))
                      ? 0 : 1;
  if (a->rowsum != nullptr) kp.direct_epi = 1;
  if (a->accumulate) kp.direct_epi = 1;   // only the register-direct epilogue reads the old C (fp32-accurate path)
  if (epilogue_override() != 0 && !a->accumulate && a->rowsum == nullptr) kp.direct_epi = epilogue_override() == 1 ? 1 : 0;
  // TMA-store epilogue (the staged groups leave through cp.async.bulk.tensor instead of ld.shared + st.global): every
  // launch whose outputs are plain row-major matrices.  CSMOE_GEMM_EPI=tma forces it where legal, =staged / =direct
  // switch it off.
  const bool tma_legal = a->c_rows == nullptr && a->c != nullptr && !a->accumulate && a->rowsum == nullptr &&
                         (kp.epi == kEpiPlain || (kp.epi == kEpiGluFwd && kp.glu_f % 32 == 0)) &&
                         (a->preact == nullptr || (a->ldpre % (a->c_dtype == CSMOE_F32 ? 4 : 8) == 0 &&
                                                   reinterpret_cast<uintptr_t>(a->preact) % 16 == 0));
  if (tma_legal && (epilogue_override() == 3 || (epilogue_override() == 0 && tma_default(a->mode)))) {
    kp.direct_epi = 0;
    kp.tma_epi = 1;
  }
  kp.stages = kPairStages;
  kp.stg_bytes = kStageTileBytes;
  // backward epilogues on the pair kernel: saved pre-activation through TMA loads (CSMOE_GEMM_BWD_TMA=0 switches it off)
  static const bool bwd_tma = []() { const char* v = getenv("CSMOE_GEMM_BWD_TMA"); return v == nullptr || v[0] != '0'; }();
  const bool bwd_tma_legal = bwd_tma && act_bwd && pair && a->c != nullptr && a->c_rows == nullptr &&
                             (kp.epi != kEpiGluBwd || kp.glu_f % 32 == 0) &&
                             reinterpret_cast<uintptr_t>(a->aux) % 16 == 0 && epilogue_override() != 1 &&
                             epilogue_override() != 2;

  Maps maps{};
  CUtensorMap& ma = maps.a;
  CUtensorMap& mb = maps.b;
  int rc;
  if (a->mode == CSMOE_GEMM_ROWS) {
    const long long a_rows = a->dense ? (a->a_expert_rows ? (long long)E * a->a_expert_rows : a->dense_rows) : a->m;
    kp.num_m_blocks = a->dense ? (kp.kcat ? 1 : E) * kp.dense_mblocks : static_cast<int>(a->m / kBM);
    kp.num_kb = static_cast<int>((a->k + kBK - 1) / kBK);
    kp.total_tiles = static_cast<long long>(kp.num_m_blocks) * kp.num_n_blocks;
    {
      cuuint64_t dims[2] = {(cuuint64_t)a->k, (cuuint64_t)a_rows};
      cuuint64_t str[1] = {(cuuint64_t)a->lda * 2};
      cuuint32_t box[2] = {kBK, kBM};   // (a cluster of two pairs loads the 128 rows as two multicast halves)
      cuuint32_t box_q[2] = {kBK, kBM / 2};
      if ((rc = encode_bf16_map(&ma, a->a, 2, dims, str, box)) != CSMOE_OK) return rc;
      if (quad && (rc = encode_bf16_map(&maps.aq, a->a, 2, dims, str, box_q)) != CSMOE_OK) return rc;
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
  if (bwd_tma_legal) {
    // (wide launches never carry a backward epilogue: see the dispatch rule below)
    const long long c_rows_total = static_cast<long long>(kp.num_m_blocks) * kBM;
    const long long cols = kp.epi == kEpiGluBwd ? 2LL * a->n : a->n;
    if ((rc = encode_out_map(&maps.c, a->c, false, cols, c_rows_total, 1, a->ldc, 0)) != CSMOE_OK) return rc;
    if ((rc = encode_out_map(&maps.z, a->aux, false, cols, c_rows_total, 1, a->ldaux, 0)) != CSMOE_OK) return rc;
    kp.tma_epi = 2;
    kp.direct_epi = 0;
    kp.stages = kPairStages - 1;
    kp.stg_bytes = 2 * kStageTileBytes;
  } else if (kp.tma_epi) {
    const bool f32 = a->c_dtype == CSMOE_F32;
    if (a->mode == CSMOE_GEMM_ROWS) {
      const long long c_rows_total = static_cast<long long>(kp.num_m_blocks) * kBM;
      if ((rc = encode_out_map(&maps.c, a->c, f32, glu_fwd ? n_grid : a->n, c_rows_total, 1, a->ldc, 0)) != CSMOE_OK) return rc;
      if (a->preact != nullptr &&
          (rc = encode_out_map(&maps.p, a->preact, f32, a->n, c_rows_total, 1, a->ldpre, 0)) != CSMOE_OK)
        return rc;
    } else {
      if ((rc = encode_out_map(&maps.c, a->c, f32, a->n, a->m, E, a->ldc, a->c_expert_stride)) != CSMOE_OK) return rc;
    }
  }
  if (pair) {
    kp.num_m_pairs = a->mode == CSMOE_GEMM_ROWS ? kp.num_m_blocks / 2 : static_cast<int>((a->m + 255) / 256);
    // 256 x 512 tiles when the output is wide enough that the second half is (almost) never padding
    // ... and the k loop is long enough to amortise the exposed epilogue (measured on B200, gemm_bench.py: k = 8192
    // +12 %, k = 16384 +8 %, k = 3072 -2 %; wgrad, whose k loop is one expert's rows, loses 10 %).  CSMOE_GEMM_WIDE
    // bits 2 / 3 force it for every ROWS / REDUCE launch.
    const int wm = wide_mask();
    const long long k_loop = a->sum_experts ? a->k * E : a->k;
    // REDUCE (wgrad): the exposed epilogue is paid back by the 25 % lower operand traffic only when the launch has many
    // waves of 256 x 512 tiles (wgrad of fc1 at the bench shape: +9 %; of fc2, 10 waves: -6 %).
    const long long wide_tiles = static_cast<long long>(kp.num_m_pairs) * ((a->n + 511) / 512) * E;
    const bool long_k = a->mode == CSMOE_GEMM_ROWS ? k_loop >= 4096 || (wm & 4)
                                                   : (wm & 8) != 0 || wide_tiles >= 16LL * (num_sms() / 2);
    // (n = 1152 as 2.25 wide blocks was tried: 0.32 vs 0.22 ms for the SigLIP fc2 -- few tiles, ragged waves)
    const bool wide = n_grid >= 512 && (n_grid % 512 == 0 || n_grid >= 2048) && long_k && kp.tma_epi != 2 &&
                      (wm & (a->mode == CSMOE_GEMM_ROWS ? 1 : 2)) != 0;
    if (wide) kp.num_n_blocks = glu_fwd ? static_cast<int>((n_grid + 255) / 256) : static_cast<int>((a->n + 511) / 512);
#if CSMOE_BUILD_QUAD
    if (quad && !wide) {
      // two pairs per cluster on neighbouring n-blocks: the cluster iterates over ceil(n-blocks / 2) groups
      const long long per_q = static_cast<long long>(kp.num_m_pairs) * ((kp.num_n_blocks + 1) / 2);
      kp.total_tiles = a->mode == CSMOE_GEMM_ROWS ? per_q : per_q * E;
      if (a->mode == CSMOE_GEMM_ROWS) maps.a = maps.aq;
      int cap = 0, rcq;
#define CSMOE_QUAD(CL_) (a->mode == CSMOE_GEMM_ROWS                                                                        \
                             ? (a->b_layout == 0 ? launch_quad<CSMOE_GEMM_ROWS, false>(maps, kp, CL_, stream, &cap)       \
                                                 : launch_quad<CSMOE_GEMM_ROWS, true>(maps, kp, CL_, stream, &cap))       \
                             : launch_quad<CSMOE_GEMM_REDUCE, true>(maps, kp, CL_, stream, &cap))
      if ((rcq = CSMOE_QUAD(0)) != CSMOE_OK) return rcq;     // occupancy query (cached after the first call)
      int clusters_q = cap;
      if (clusters_q <= 0) {
        set_error("csmoe_grouped_gemm: no cluster of 4 CTAs fits on this device");
        return CSMOE_ERR_CUDA;
      }
      if (a->max_ctas > 3 && a->max_ctas / 4 < clusters_q) clusters_q = a->max_ctas / 4;
      if (kp.total_tiles < clusters_q) clusters_q = static_cast<int>(kp.total_tiles);
      kp.stats = stats_buffer();
      if (kp.stats != nullptr) cudaMemsetAsync(kp.stats, 0, 256 * 8 * sizeof(unsigned long long), stream);
      rcq = CSMOE_QUAD(clusters_q);
#undef CSMOE_QUAD
      if (rcq == CSMOE_OK && kp.stats != nullptr) {
        char what[96];
        snprintf(what, sizeof(what), "quad(%d clusters) mode=%d b_layout=%d epi=%d n=%lld k=%lld", clusters_q, a->mode,
                 a->b_layout, kp.epi, (long long)a->n, (long long)a->k);
        print_stats(what, kp, 4 * clusters_q, stream);
      }
      return rcq;
    }
#endif
    const long long per = static_cast<long long>(kp.num_m_pairs) * kp.num_n_blocks;
    kp.total_tiles = a->mode == CSMOE_GEMM_ROWS ? per : per * E;
    int clusters = num_sms() / 2;
    if (clusters <= 0) return CSMOE_ERR_CUDA;
    if (a->max_ctas > 1 && a->max_ctas / 2 < clusters) clusters = a->max_ctas / 2;
    if (kp.total_tiles < clusters) clusters = static_cast<int>(kp.total_tiles);
    kp.stats = stats_buffer();
    if (kp.stats != nullptr) {
      cudaMemsetAsync(kp.stats, 0, 256 * 8 * sizeof(unsigned long long), stream);
      int rc2;
      if (wide) {
        if (a->mode == CSMOE_GEMM_ROWS)
          rc2 = a->b_layout == 0 ? launch_wide<CSMOE_GEMM_ROWS, false>(maps, kp, clusters, stream)
                                 : launch_wide<CSMOE_GEMM_ROWS, true>(maps, kp, clusters, stream);
        else
          rc2 = launch_wide<CSMOE_GEMM_REDUCE, true>(maps, kp, clusters, stream);
      } else if (a->mode == CSMOE_GEMM_ROWS) {
        rc2 = a->b_layout == 0 ? launch_pair<CSMOE_GEMM_ROWS, false>(maps, kp, clusters, stream)
                               : launch_pair<CSMOE_GEMM_ROWS, true>(maps, kp, clusters, stream);
      } else {
        rc2 = launch_pair<CSMOE_GEMM_REDUCE, true>(maps, kp, clusters, stream);
      }
      if (rc2 == CSMOE_OK) {
        char what[96];
        snprintf(what, sizeof(what), "%s mode=%d b_layout=%d epi=%d n=%lld k=%lld", wide ? "wide" : "pair", a->mode,
                 a->b_layout, kp.epi, (long long)a->n, (long long)a->k);
        print_stats(what, kp, 2 * clusters, stream);
      }
      return rc2;
    }
    if (wide) {
      if (a->mode == CSMOE_GEMM_ROWS)
        return a->b_layout == 0 ? launch_wide<CSMOE_GEMM_ROWS, false>(maps, kp, clusters, stream)
                                : launch_wide<CSMOE_GEMM_ROWS, true>(maps, kp, clusters, stream);
      return launch_wide<CSMOE_GEMM_REDUCE, true>(maps, kp, clusters, stream);
    }
    if (a->mode == CSMOE_GEMM_ROWS)
      return a->b_layout == 0 ? launch_pair<CSMOE_GEMM_ROWS, false>(maps, kp, clusters, stream)
                              : launch_pair<CSMOE_GEMM_ROWS, true>(maps, kp, clusters, stream);
    return launch_pair<CSMOE_GEMM_REDUCE, true>(maps, kp, clusters, stream);
  }
  int grid = num_sms();
  if (grid <= 0) return CSMOE_ERR_CUDA;
  if (a->max_ctas > 0 && a->max_ctas < grid) grid = a->max_ctas;
  if (kp.total_tiles < grid) grid = static_cast<int>(kp.total_tiles);

  if (a->mode == CSMOE_GEMM_ROWS) {
    if (a->b_layout == 0)
      return big_n ? launch<CSMOE_GEMM_ROWS, false, 256>(maps, kp, grid, stream)
                   : launch<CSMOE_GEMM_ROWS, false, 128>(maps, kp, grid, stream);
    return big_n ? launch<CSMOE_GEMM_ROWS, true, 256>(maps, kp, grid, stream)
                 : launch<CSMOE_GEMM_ROWS, true, 128>(maps, kp, grid, stream);
  }
  return big_n ? launch<CSMOE_GEMM_REDUCE, true, 256>(maps, kp, grid, stream)
               : launch<CSMOE_GEMM_REDUCE, true, 128>(maps, kp, grid, stream);
}
