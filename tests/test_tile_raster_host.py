"""The grouped GEMM's tile rasters, checked on the CPU from the kernel's OWN source text.

`decode_tile` / `decode_tile_pair` (competesmoe_b200/csrc/gemm_tcgen05.cu) map a persistent CTA's tile counter to an
output tile (expert, m-block, n-block).  They are pure integer code, so this test cuts `struct KParams`, `struct Tile`
and the two functions out of the .cu file as they are, compiles them for the host with g++ (`__device__` defined away,
`__ldg` as a plain load) and enumerates every tile counter of a launch: whatever the raster (n-fastest bands, m-fastest
bands, bands aligned to the experts' row ranges), band width, band cap, ragged / empty experts and unused tail of the
statically sized row space, every output tile that holds routed rows must be produced exactly once, with the expert the
row map names, and nothing else may be produced.  A raster that skips or repeats a tile gives wrong results on the GPU;
this catches it without one.  (The GPU tests run the same launches under every raster and compare results bit for bit.)
"""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from host_extract import block, compile_host

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "competesmoe_b200" / "csrc" / "gemm_tcgen05.cu"

PRELUDE = r"""
#include <algorithm>
#include <cstdint>
#include <cstring>
#include "csmoe.h"
#define __device__
#define __forceinline__ inline
template <class T> static inline T __ldg(const T* p) { return *p; }
using std::min;
constexpr int kBM = 128, kBK = 64;
"""

HARNESS = r"""
// counts[(e * units_m + mb) * num_n + nb] += 1 for every valid tile; returns the number of tiles flagged invalid, or a
// negative code on the first inconsistency.
template <int MODE, bool PAIR>
static long long walk(const KParams& p, int units_m, int num_n, int* counts, int rank) {
  long long invalid = 0;
  for (long long t = 0; t < p.total_tiles; ++t) {
    const Tile ti = PAIR ? decode_tile_pair<MODE>(p, t, rank) : decode_tile<MODE>(p, t);
    if (ti.mb < 0 || ti.mb >= units_m || ti.nb < 0 || ti.nb >= num_n) return -1;
    if (!ti.valid) {
      if (MODE != CSMOE_GEMM_ROWS) return -2;                      // REDUCE tiles are always valid
      if (p.dense || p.kcat) return -3;
      if (p.tile_expert[PAIR ? 2 * ti.mb : ti.mb] >= 0) return -4;  // a routed row tile was skipped
      ++invalid;
      continue;
    }
    if (ti.e < 0 || ti.e >= p.num_experts) return -5;
    const int row_unit = PAIR ? 256 : 128;
    if (MODE == CSMOE_GEMM_ROWS) {
      if (!p.dense && !p.kcat) {
        if (p.tile_expert[PAIR ? 2 * ti.mb : ti.mb] != ti.e) return -6;
        if (ti.a_row != ti.mb * row_unit + (PAIR ? rank * kBM : 0)) return -7;
      }
      if (p.dense && !p.kcat) {
        const int per_e = PAIR ? p.dense_mblocks / 2 : p.dense_mblocks;
        if (ti.e != ti.mb / per_e) return -8;
        if (ti.a_row != (ti.mb % per_e) * row_unit + (PAIR ? rank * kBM : 0) + ti.e * p.a_expert_rows) return -9;
      }
      if (ti.nkb != (p.kcat ? p.num_kb * p.num_experts : p.num_kb)) return -10;
      counts[static_cast<long long>(ti.mb) * num_n + ti.nb] += 1;
    } else {
      if (!p.dense) {
        if (ti.a_row != p.pad_offsets[ti.e] || ti.b_row != ti.a_row) return -11;
        if (ti.nkb != (p.pad_offsets[ti.e + 1] - p.pad_offsets[ti.e]) / kBK) return -12;
      }
      counts[(static_cast<long long>(ti.e) * units_m + ti.mb) * num_n + ti.nb] += 1;
    }
  }
  return invalid;
}

extern "C" long long raster_walk(int mode, int pair, int raster_m, int band, int band_cap, int units_m, int num_n,
                                 int num_experts, int dense, int dense_mblocks, int a_expert_rows, int kcat,
                                 const int* tile_expert, const int* pad_offsets, int* counts, int rank) {
  KParams p;
  memset(&p, 0, sizeof(p));
  p.num_experts = num_experts;
  p.num_n_blocks = num_n;
  p.num_kb = 3;
  p.band = band;
  p.raster_m = raster_m;
  p.band_cap = band_cap;
  p.dense = dense;
  p.dense_mblocks = dense_mblocks;
  p.dense_kblocks = dense_mblocks * 2;
  p.a_expert_rows = a_expert_rows;
  p.b_expert_rows = a_expert_rows;
  p.kcat = kcat;
  p.tile_expert = tile_expert;
  p.pad_offsets = pad_offsets;
  if (pair) {
    p.num_m_pairs = units_m;
    p.num_m_blocks = 2 * units_m;
  } else {
    p.num_m_blocks = units_m;
  }
  const long long per = static_cast<long long>(units_m) * num_n;
  p.total_tiles = mode == CSMOE_GEMM_ROWS ? per : per * num_experts;
  if (mode == CSMOE_GEMM_ROWS) return pair ? walk<CSMOE_GEMM_ROWS, true>(p, units_m, num_n, counts, rank)
                                           : walk<CSMOE_GEMM_ROWS, false>(p, units_m, num_n, counts, rank);
  return pair ? walk<CSMOE_GEMM_REDUCE, true>(p, units_m, num_n, counts, rank)
              : walk<CSMOE_GEMM_REDUCE, false>(p, units_m, num_n, counts, rank);
}
"""


@pytest.fixture(scope="module")
def raster(tmp_path_factory):
    text = SRC.read_text()
    parts = [block(text, r"struct KParams\s*\{", SRC.name), block(text, r"struct Tile\s*\{", SRC.name),
             block(text, r"template <int MODE>\s*__device__ __forceinline__ Tile decode_tile\(", SRC.name),
             block(text, r"template <int MODE>\s*__device__ __forceinline__ Tile decode_tile_pair\(", SRC.name)]
    lib = compile_host(PRELUDE + "\n".join(parts) + HARNESS, tmp_path_factory.mktemp("raster"), "raster_host")
    lib.raster_walk.restype = C.c_longlong
    lib.raster_walk.argtypes = [C.c_int] * 12 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return lib


ROWS, REDUCE = 0, 1


def _route(counts, row_tile, spare_tiles):
    """pad_offsets / tile_expert of ops.route_build for per-expert row counts (segments aligned to row_tile, a tail of
    unused 128-row tiles marked -1: the row space is sized for the worst case)."""
    pad = [0]
    for c in counts:
        pad.append(pad[-1] + (c + row_tile - 1) // row_tile * row_tile)
    cap = pad[-1] + spare_tiles * row_tile
    te = np.full(cap // 128, -1, dtype=np.int32)
    for e in range(len(counts)):
        te[pad[e] // 128:pad[e + 1] // 128] = e
    return np.asarray(pad, dtype=np.int32), te, cap


def _walk(lib, mode, pair, raster_m, band, cap, units_m, num_n, E, te=None, po=None, dense=0, dense_mblocks=0, a_rows=0,
          kcat=0, rank=0):
    n = units_m * num_n * (E if mode == REDUCE else 1)
    counts = np.zeros(max(n, 1), dtype=np.int32)
    rc = lib.raster_walk(mode, int(pair), raster_m, band, cap, units_m, num_n, E, dense, dense_mblocks, a_rows, kcat,
                         None if te is None else te.ctypes.data, None if po is None else po.ctypes.data,
                         counts.ctypes.data, rank)
    return rc, counts[:n]


@pytest.mark.parametrize("seed", range(6))
def test_routed_rows_every_raster_is_a_bijection(raster, seed):
    rng = np.random.default_rng(seed)
    E = int(rng.choice([1, 2, 4, 7, 16, 64]))
    counts = rng.integers(0, 2600, size=E)
    counts[rng.integers(0, E)] = 0                                   # an empty expert
    if E > 1:
        counts[rng.integers(0, E)] = int(rng.integers(2000, 9000))   # and a hot one
    for row_tile, pair in ((256, True), (256, False), (128, False)):
        po, te, cap = _route(counts.tolist(), row_tile, spare_tiles=int(rng.integers(0, 5)))
        units = cap // (256 if pair else 128)
        for num_n in (1, 2, 9, 12, 64):
            for raster_m, band, bcap in ((0, 8, 12), (0, 3, 12), (1, 8, 12), (1, 5, 12), (2, 8, 12), (2, 8, 1), (2, 8, 2), (2, 8, 5)):
                if not pair and raster_m != 0:
                    continue                                         # the single-CTA kernel has the n-fastest raster only
                for rank in ((0, 1) if pair else (0,)):
                    rc, got = _walk(raster, ROWS, pair, raster_m, band, bcap, units, num_n, E, te, po, rank=rank)
                    used = te[::2] >= 0 if pair else te >= 0
                    assert rc == int((~used).sum()) * num_n, \
                        f"rc {rc}: seed {seed} E {E} row_tile {row_tile} pair {pair} n-blocks {num_n} raster {raster_m}/{band}/{bcap}"
                    want = np.repeat(used.astype(np.int32), num_n)
                    assert np.array_equal(got, want), \
                        f"tiles skipped or repeated: seed {seed} E {E} pair {pair} n-blocks {num_n} raster {raster_m}/{band}/{bcap}"


def test_bench_shape_expert_bands(raster):
    # configs[1]: 4096 tokens x top-2 over 4 experts, 256-row segments, fc1 with n = 16384 as 64 n-blocks of 256
    po, te, cap = _route([2048 + 45, 2048 - 45, 2048 + 7, 2048 - 7], 256, spare_tiles=0)
    rc, got = _walk(raster, ROWS, True, 2, 8, 12, cap // 256, 64, 4, te, po)
    assert rc == 0 and bool((got == 1).all())


@pytest.mark.parametrize("pair", [False, True])
def test_dense_and_reduce_launches(raster, pair):
    unit = 256 if pair else 128
    for E in (1, 4, 8):
        for t_pad in (256, 1024, 4352):
            per_e = t_pad // unit
            for num_n in (1, 3, 16):
                for raster_m, band in ((0, 8), (1, 8), (0, 2)):
                    if not pair and raster_m:
                        continue
                    # competition step: every expert x the same rows (ROWS, dense)
                    rc, got = _walk(raster, ROWS, pair, raster_m, band, 12, E * per_e, num_n, E, dense=1,
                                    dense_mblocks=t_pad // 128, a_rows=t_pad)
                    assert rc == 0 and bool((got == 1).all()), (E, t_pad, num_n, raster_m, band)
                    # its dx: one output block per row block, the k loop runs over the experts (sum_experts)
                    rc, got = _walk(raster, ROWS, pair, raster_m, band, 12, per_e, num_n, E, dense=1,
                                    dense_mblocks=t_pad // 128, a_rows=t_pad, kcat=1)
                    assert rc == 0 and bool((got == 1).all())
                # weight gradients: per expert an [m, n] block, contraction over that expert's rows
                po, te, cap = _route([300 * (e + 1) for e in range(E)], 256, 1)
                for units_m in (1, 2, 5):
                    for band in (8, 3):
                        rc, got = _walk(raster, REDUCE, pair, 0, band, 12, units_m, num_n, E, te, po)
                        assert rc == 0 and bool((got == 1).all()), (E, units_m, num_n, band)
                        rc, got = _walk(raster, REDUCE, pair, 0, band, 12, units_m, num_n, E, dense=1,
                                        dense_mblocks=t_pad // 128, a_rows=t_pad)
                        assert rc == 0 and bool((got == 1).all())
