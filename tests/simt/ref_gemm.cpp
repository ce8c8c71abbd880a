// TEST INFRASTRUCTURE ONLY -- NOT the tensor-core kernel and no evidence about it.
//
// A plain-loop statement of csmoe_grouped_gemm's CONTRACT (include/csmoe.h; rounding points of the epilogues as in
// csrc/gemm_tcgen05.cu: epilogue_store8 / epilogue_tile) so that the layers' host logic and every non-GEMM kernel can run
// end to end on the SIMT emulator (tests/simt/simt.h) against the golden fixtures of the reference.  The tcgen05 / TMA
// kernels themselves are only ever tested on the GPU (`-m gpu`).  fp32 accumulation in ascending k order; the scalar
// activation functions are the shipped ones (common.h).
#include <vector>

#include "common.h"

using namespace csmoe;

namespace {

inline float bf(const void* p, long long i) { return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]); }
inline float ld_out(const void* p, long long i, bool f32) { return f32 ? reinterpret_cast<const float*>(p)[i] : bf(p, i); }
inline void st_out(void* p, long long i, bool f32, float v) {
  if (f32) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}
inline float rnd(float v, bool f32) { return f32 ? v : bf16_round(v); }
inline float softplus_ref(float x) { return fmaxf(x, 0.f) + logf(1.f + expf(-fabsf(x))); }

// B[e] as a dense fp32 [k][n] matrix
std::vector<float> expert_b(const csmoe_gemm_args* a, int e) {
  std::vector<float> w(static_cast<size_t>(a->k) * a->n);
  const long long base = static_cast<long long>(e) * a->b_expert_stride;
  for (long long kk = 0; kk < a->k; ++kk)
    for (long long j = 0; j < a->n; ++j)
      w[kk * a->n + j] = bf(a->b, base + (a->b_layout == 0 ? j * a->ldb + kk : kk * a->ldb + j));
  return w;
}

int gemm_rows(const csmoe_gemm_args* a) {
  const long long n = a->n, k = a->k;
  const bool f32 = a->c_dtype == CSMOE_F32;
  const bool glu_fwd = a->act == CSMOE_ACT_SILU_GLU;
  const bool act_bwd = a->act_bwd != CSMOE_ACT_NONE;
  const bool glu_bwd = a->act_bwd == CSMOE_ACT_SILU_GLU;
  const long long F = glu_fwd ? n / 2 : n;
  std::vector<std::vector<float>> W(a->num_experts);
  std::vector<float> acc(n), arow(k);
  const long long groups = (n + 63) / 64;
  for (long long r = 0; r < a->m; ++r) {
    std::fill(acc.begin(), acc.end(), 0.f);
    int e_bias = 0;
    auto add = [&](int e, long long a_row) {
      if (W[e].empty()) W[e] = expert_b(a, e);
      for (long long kk = 0; kk < k; ++kk) arow[kk] = bf(a->a, a_row * a->lda + kk);
      const float* w = W[e].data();
      for (long long kk = 0; kk < k; ++kk) {
        const float av = arow[kk];
        if (av == 0.f) continue;
        const float* wr = w + kk * n;
        for (long long j = 0; j < n; ++j) acc[j] += av * wr[j];
      }
    };
    if (a->sum_experts) {
      for (int e = 0; e < a->num_experts; ++e) add(e, e * a->a_expert_rows + r);
    } else if (a->dense) {
      const int e = static_cast<int>(r / a->dense_rows);
      add(e, (r % a->dense_rows) + e * a->a_expert_rows);
      e_bias = e;
    } else {
      const int e = a->tile_expert[r / 128];
      if (e < 0) continue;                       // row tile past the routed rows: outputs left untouched
      add(e, r);
      e_bias = e;
    }
    void* crow = a->c;
    long long cbase = r * a->ldc;
    if (a->c_rows != nullptr) {
      if (a->c_rows[r] == 0) continue;
      crow = reinterpret_cast<void*>(a->c_rows[r]);
      cbase = 0;
    }
    if (glu_fwd) {
      for (long long f = 0; f < F; ++f) {
        const float zg = bf16_round(acc[f]), zu = bf16_round(acc[F + f]);
        st_out(a->preact, r * a->ldpre + f, false, zg);
        st_out(a->preact, r * a->ldpre + F + f, false, zu);
        st_out(crow, cbase + f, false, zu * bf16_round(act_apply(zg, CSMOE_ACT_SILU, true)));
      }
      continue;
    }
    if (act_bwd) {
      for (long long j = 0; j < n; ++j) {
        const float dh = bf16_round(acc[j]);
        const float z0 = bf(a->aux, r * a->ldaux + j);
        if (glu_bwd) {
          const float z1 = bf(a->aux, r * a->ldaux + n + j);
          const float sg = bf16_round(act_apply(z0, CSMOE_ACT_SILU, true));
          st_out(crow, cbase + n + j, false, dh * sg);
          st_out(crow, cbase + j, false, bf16_round(dh * z1) * act_grad(z0, CSMOE_ACT_SILU, true));
        } else {
          st_out(crow, cbase + j, false, dh * act_grad(z0, a->act_bwd, true));
        }
      }
      continue;
    }
    std::vector<float> rs(groups, 0.f);
    for (long long j = 0; j < n; ++j) {
      float z = acc[j];
      if (a->bias != nullptr) {
        const long long bi = static_cast<long long>(e_bias) * n + j;
        const float b = a->bias_dtype == CSMOE_F32 ? reinterpret_cast<const float*>(a->bias)[bi] : bf(a->bias, bi);
        z = (a->bias_after_round ? bf16_round(z) : z) + b;
      }
      if (a->accumulate) z += ld_out(crow, cbase + j, f32);
      if (a->act != CSMOE_ACT_NONE || a->preact != nullptr) {
        z = rnd(z, f32);
        if (a->preact != nullptr) st_out(a->preact, r * a->ldpre + j, f32, z);
        z = act_apply(z, a->act, !f32);
      }
      st_out(crow, cbase + j, f32, z);
      if (a->rowsum != nullptr) {
        const float sp = softplus_ref(rnd(z, f32));
        rs[j / 64] += a->rowsum_round ? bf16_round(sp) : sp;
      }
    }
    if (a->rowsum != nullptr)
      for (long long g = 0; g < groups; ++g) a->rowsum[r * groups + g] = rs[g];
  }
  return CSMOE_OK;
}

int gemm_reduce(const csmoe_gemm_args* a) {
  const bool f32 = a->c_dtype == CSMOE_F32;
  const long long m = a->m, n = a->n;
  std::vector<float> acc(static_cast<size_t>(m) * n), ar(m), br(n);
  for (int e = 0; e < a->num_experts; ++e) {
    long long a0, b0, rows;
    if (a->dense) {
      a0 = static_cast<long long>(e) * a->a_expert_rows;
      b0 = static_cast<long long>(e) * a->b_expert_stride;
      rows = a->dense_rows;
    } else {
      a0 = b0 = a->pad_offsets[e];
      rows = a->pad_offsets[e + 1] - a->pad_offsets[e];
    }
    std::fill(acc.begin(), acc.end(), 0.f);
    for (long long r = 0; r < rows; ++r) {
      for (long long i = 0; i < m; ++i) ar[i] = bf(a->a, (a0 + r) * a->lda + i);
      for (long long j = 0; j < n; ++j) br[j] = bf(a->b, (b0 + r) * a->ldb + j);
      for (long long i = 0; i < m; ++i) {
        const float av = ar[i];
        if (av == 0.f) continue;
        float* row = acc.data() + i * n;
        for (long long j = 0; j < n; ++j) row[j] += av * br[j];
      }
    }
    const long long cb = static_cast<long long>(e) * a->c_expert_stride;
    for (long long i = 0; i < m; ++i)
      for (long long j = 0; j < n; ++j) {
        float z = acc[i * n + j];
        if (a->accumulate) z += ld_out(a->c, cb + i * a->ldc + j, f32);
        st_out(a->c, cb + i * a->ldc + j, f32, z);
      }
  }
  return CSMOE_OK;
}

}  // namespace

extern "C" int csmoe_grouped_gemm(const csmoe_gemm_args* a, void*) {
  CSMOE_CHECK_ARG(a != nullptr, "csmoe_grouped_gemm: args is NULL");
  CSMOE_CHECK_ARG(a->a && a->b && (a->c || a->c_rows), "csmoe_grouped_gemm: a/b/c must be non-NULL");
  CSMOE_CHECK_ARG(a->m > 0 && a->n > 0 && a->k > 0 && a->n % 8 == 0, "csmoe_grouped_gemm: bad sizes");
  if (a->mode == CSMOE_GEMM_ROWS) {
    CSMOE_CHECK_ARG(a->dense || a->tile_expert != nullptr, "csmoe_grouped_gemm: ROWS mode needs tile_expert");
    return gemm_rows(a);
  }
  CSMOE_CHECK_ARG(a->dense || a->pad_offsets != nullptr, "csmoe_grouped_gemm: REDUCE mode needs pad_offsets");
  return gemm_reduce(a);
}

// the fused sigma-MoE kernels are tensor-core kernels: never available on the emulator (callers take the grouped GEMMs)
extern "C" int csmoe_sigma_ffn_supported(int64_t, int32_t, int64_t) { return 0; }
