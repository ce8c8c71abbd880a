"""Scalar device functions of the shipped kernels, compiled for the host from their own source text (tests/host_extract.py)
and checked on the CPU:

* `philox4x32_10` / `keep8` (csrc/block.cu): the counter-based stream behind the fused residual + dropout epilogue.
  Known-answer vectors of Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the
  Random123 distribution's kat_vectors), an independent numpy statement of the round function, the (seed, element index)
  -> counter mapping forward and backward rely on, and the keep rate.
* `act_apply` / `act_grad` (+ the vector forms) (csrc/common.h): ReLU / GELU / GELU-tanh / SiLU values against torch and
  their derivatives against torch autograd -- the epilogues of the grouped GEMM and the activation-backward kernels
  multiply by exactly these.
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from host_extract import CSRC, block, compile_host

CUDA_SHIMS = r"""
#include <cmath>
#include <cstdint>
#include "csmoe.h"
#define __device__
#define __forceinline__ inline
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32); }
#define __expf(x) expf(x)                  /* glibc declares a function of that name: a macro, after <cmath> */
#define __fdividef(a, b) ((a) / (b))
static inline float tanh_approx(float x) { return tanhf(x); }     // tanh.approx.f32 is inline PTX: the host uses tanhf
"""


# ------------------------------------------------------------------------------------------------ Philox
@pytest.fixture(scope="module")
def philox(tmp_path_factory):
    text = (CSRC / "block.cu").read_text()
    parts = [block(text, r"__device__ __forceinline__ uint4 philox4x32_10\(", "block.cu"),
             block(text, r"__device__ __forceinline__ void keep8\(", "block.cu")]
    harness = r"""
extern "C" void philox_host(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  const uint4 r = philox4x32_10(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]));
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
extern "C" void keep8_host(unsigned long long seed, long long e0, uint32_t threshold, uint8_t* keep) {
  bool k[8];
  keep8(seed, e0, threshold, k);
  for (int i = 0; i < 8; ++i) keep[i] = k[i];
}
"""
    lib = compile_host(CUDA_SHIMS + "\n".join(parts) + harness, tmp_path_factory.mktemp("philox"), "philox_host")
    lib.philox_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.keep8_host.argtypes = [C.c_ulonglong, C.c_longlong, C.c_uint32, C.c_void_p]
    return lib


def _philox(lib, ctr, key):
    c, k, o = np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), np.zeros(4, np.uint32)
    lib.philox_host(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return [int(v) for v in o]


def _philox_numpy(ctr, key, rounds=10):
    """Philox4x32 as the paper states it: two 32x32 -> 64 multiplies per round, key bumped by the Weyl constants."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c, k = [int(v) for v in ctr], [int(v) for v in key]
    for r in range(rounds):
        if r > 0:
            k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
    return c


KAT = [  # Random123 kat_vectors, "philox4x32 10": counter, key, expected
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


@pytest.mark.parametrize("ctr, key, want", KAT)
def test_philox_known_answers(philox, ctr, key, want):
    assert _philox_numpy(ctr, key) == want, "the numpy statement of Philox4x32-10 disagrees with the published vector"
    assert _philox(philox, ctr, key) == want


def test_philox_matches_the_independent_statement_on_random_inputs(philox):
    rng = np.random.default_rng(0)
    for _ in range(200):
        ctr, key = rng.integers(0, 2 ** 32, 4, dtype=np.uint64), rng.integers(0, 2 ** 32, 2, dtype=np.uint64)
        assert _philox(philox, ctr, key) == _philox_numpy(ctr, key)


def _keep8(lib, seed, e0, thr):
    k = np.zeros(8, np.uint8)
    lib.keep8_host(seed, e0, thr, k.ctypes.data)
    return k.astype(bool)


def test_dropout_mask_is_a_function_of_seed_and_element_index(philox):
    """keep8(seed, e0) covers elements e0 .. e0+7 with counters e0/4 and e0/4 + 1 (four 32-bit draws each), key = the
    64-bit seed: the backward kernel regenerates the same bits from the same (seed, index), whatever thread asks."""
    seed, thr = 0x1234_5678_9ABC_DEF0, int(0.1 * 2 ** 32)
    key = [seed & 0xFFFFFFFF, seed >> 32]
    for e0 in (0, 8, 4096, 8 * 123457, (1 << 34) + 16):
        blk = e0 >> 2
        draws = _philox_numpy([blk & 0xFFFFFFFF, blk >> 32, 0, 0], key) + \
            _philox_numpy([(blk + 1) & 0xFFFFFFFF, (blk + 1) >> 32, 0, 0], key)
        assert _keep8(philox, seed, e0, thr).tolist() == [d >= thr for d in draws]
    # neighbouring 8-element groups share nothing, and another seed gives another mask
    a = np.concatenate([_keep8(philox, seed, 8 * i, 1 << 31) for i in range(64)])
    b = np.concatenate([_keep8(philox, seed + 1, 8 * i, 1 << 31) for i in range(64)])
    assert 0.35 < a.mean() < 0.65 and 0.3 < (a != b).mean() < 0.7


@pytest.mark.parametrize("p", [0.0, 0.1, 0.5])
def test_dropout_keep_rate(philox, p):
    thr = min(int(p * 2 ** 32), 2 ** 32 - 1)
    keep = np.concatenate([_keep8(philox, 42, 8 * i, thr) for i in range(8192)])
    assert abs(keep.mean() - (1 - p)) < 4 * np.sqrt(max(p * (1 - p), 1e-12) / keep.size) + 1e-12


# ------------------------------------------------------------------------------------------------ activations
ACTS = {"relu": 1, "gelu": 2, "gelu_tanh": 3, "silu": 4}


@pytest.fixture(scope="module")
def acts(tmp_path_factory):
    text = (CSRC / "common.h").read_text()
    parts = [block(text, r"__device__ __forceinline__ float act_apply\(", "common.h"),
             block(text, r"__device__ __forceinline__ float act_grad\(", "common.h"),
             block(text, r"template <int N>\s*__device__ __forceinline__ void act_apply_vec\(", "common.h"),
             block(text, r"template <int N>\s*__device__ __forceinline__ void act_grad_vec\(", "common.h")]
    harness = r"""
extern "C" void act_host(const float* z, int n, int act, int fast, int vec, float* val, float* grad) {
  if (!vec) {
    for (int i = 0; i < n; ++i) { val[i] = act_apply(z[i], act, fast != 0); grad[i] = act_grad(z[i], act, fast != 0); }
    return;
  }
  for (int i = 0; i + 8 <= n; i += 8) {
    float v[8], g[8], zz[8];
    for (int j = 0; j < 8; ++j) { v[j] = z[i + j]; zz[j] = z[i + j]; g[j] = 1.f; }
    act_apply_vec<8>(v, act, fast != 0);
    act_grad_vec<8>(g, zz, act, fast != 0);
    for (int j = 0; j < 8; ++j) { val[i + j] = v[j]; grad[i + j] = g[j]; }
  }
}
"""
    lib = compile_host(CUDA_SHIMS + "\n".join(parts) + harness, tmp_path_factory.mktemp("acts"), "acts_host")
    lib.act_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def test_activation_codes_match_the_header():
    import re
    from host_extract import ROOT
    text = (ROOT / "include" / "csmoe.h").read_text()
    for name, code in (("RELU", 1), ("GELU", 2), ("GELU_TANH", 3), ("SILU", 4)):
        m = re.search(rf"CSMOE_ACT_{name}\s*=\s*(\d+)", text)
        assert m and int(m.group(1)) == code, f"CSMOE_ACT_{name} is no longer {code}: update ACTS"


@pytest.mark.parametrize("name", sorted(ACTS))
@pytest.mark.parametrize("vec", [0, 1])
@pytest.mark.parametrize("fast", [0, 1])
def test_activation_value_and_derivative_match_torch(acts, name, vec, fast):
    z = torch.cat([torch.linspace(-12, 12, 4001), torch.tensor([0.0, -0.0, 1e-6, -1e-6, 30.0, -30.0, 88.0])])
    z = z[: z.numel() // 8 * 8].contiguous()
    zt = z.clone().double().requires_grad_(True)
    fn = {"relu": F.relu, "gelu": F.gelu, "gelu_tanh": lambda t: F.gelu(t, approximate="tanh"), "silu": F.silu}[name]
    want = fn(zt)
    (dwant,) = torch.autograd.grad(want.sum(), zt)
    zn = z.numpy().astype(np.float32)
    val, grad = np.zeros_like(zn), np.zeros_like(zn)
    acts.act_host(zn.ctypes.data, zn.size, ACTS[name], fast, vec, val.ctypes.data, grad.ctypes.data)
    if name == "relu":
        dwant = (z > 0).double()             # the kernels define relu'(0) = 0 (torch agrees)
    np.testing.assert_allclose(val, want.detach().numpy(), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(grad, dwant.numpy(), rtol=4e-6, atol=4e-6)
