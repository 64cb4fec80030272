// tile_spmv.cuh -- the SpMV engine of the persistent kernels: y_row = sum_k val[k] * vec[idx[k]] over
// statically tiled CSR rows, one call per phase, the row result handed to a fused epilogue.
//
// Why this shape (tools/gather_bench.cu, measured on B200): a random fp64 gather out of L2 costs one L1TEX
// wavefront, and an SM retires ~1 wavefront per clock -- but only if every load instruction is fully populated
// and enough of them are in flight.  Lanes-per-row CSR kernels reach half of that on 10-nnz rows.  So:
//   * the matrix stream (values + column indices of a tile of <= TS_TILE nonzeros) is staged into shared
//     memory by TMA bulk copies (cp.async.bulk + mbarrier complete_tx, TS_STAGES-deep ring, L2 evict_first):
//     it never passes through the L1TEX tag stage and its HBM latency is hidden by the ring;
//   * gathers are issued nonzero-parallel (thread t takes nonzeros t, t+NT, ...), TS_J independent loads per
//     thread, one tile AHEAD of the arithmetic: the gathers of tile i+1 are in flight while the rows of tile i
//     are summed;
//   * products go back into the value slot of the stage; rows are then summed out of shared memory in storage
//     order (one thread per row, or 2..32 lanes per row when a tile has few rows), so results are bitwise
//     reproducible; a row longer than a tile is cut into chunk tiles whose partial sums are carried by thread 0.
// Reference counterpart: the SpMVs of the indirect KKT solve.
#pragma once
#include "device_common.cuh"

namespace b200qp {

#ifndef B200QP_TS_TILE
#define B200QP_TS_TILE 2048
#endif
#ifndef B200QP_TS_STAGES
#define B200QP_TS_STAGES 4
#endif
constexpr int TS_TILE = B200QP_TS_TILE;       // nonzeros per tile
constexpr int TS_CAP = TS_TILE + 8;           // stage capacity (16-byte alignment slack on both ends)
constexpr int TS_STAGES = B200QP_TS_STAGES;
constexpr int TS_J = TS_TILE / NT;            // gathers per thread per tile
constexpr int TS_MAXROWS = 4 * NT;            // rows per tile
constexpr size_t TS_SMEM = (size_t)TS_STAGES * TS_CAP * 12;
static_assert(TS_TILE % NT == 0, "tile must be a multiple of the CTA size");

// tile descriptor: x = first row, y = one past the last row (whole-row tile) or -1 / -2 (chunk of the long row x;
// -2 = last chunk), z..w = nonzero range
struct TileMat {
  int rows, cols, nnz, ntiles;
  const int* __restrict__ ptr;
  const int* __restrict__ idx;     // padded by 8 entries
  const double* __restrict__ val;  // padded by 8 entries
  const int4* __restrict__ tiles;
  const int* __restrict__ split;   // merged matrices [M1 | M2]: first nonzero of the M2 part of every row (or null)
};

constexpr int TS_RING = 16;        // descriptor ring (shared memory), >= 2 * TS_STAGES

struct TileCtx {                   // per-CTA pipeline state, lives across all phases of a kernel
  double* sval;                    // [TS_STAGES][TS_CAP]
  int* sidx;                       // [TS_STAGES][TS_CAP]
  uint64_t* full;                  // [TS_STAGES] mbarriers
  double* s_part;                  // [2 * NW] scratch for chunk tiles
  int4* s_td;                      // [TS_RING] tile descriptors
  int stage;                       // stage of the next tile to consume (identical in every thread)
  unsigned parity;                 // bit s = parity to wait for on stage s
  uint64_t policy;
};

// shared memory a kernel using the engine must provide besides the TS_SMEM dynamic bytes
struct TileShared {
  alignas(16) int4 td[TS_RING];
  alignas(8) uint64_t bars[TS_STAGES];
  double part[2 * NW];
};

__device__ __forceinline__ void tile_ctx_init(TileCtx& cx, unsigned char* smem_raw, TileShared& sh) {
  cx.sval = reinterpret_cast<double*>(smem_raw);
  cx.sidx = reinterpret_cast<int*>(smem_raw + (size_t)TS_STAGES * TS_CAP * 8);
  cx.full = sh.bars;
  cx.s_part = sh.part;
  cx.s_td = sh.td;
  cx.stage = 0;
  cx.parity = 0u;
  cx.policy = policy_evict_first();
  if (threadIdx.x == 0) {
    for (int s = 0; s < TS_STAGES; ++s) mbar_init(&sh.bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
}

// SPLIT = false: epi(row, sum, 0)      SPLIT = true: epi(row, sum of the M1 part, sum of the M2 part)
// Exactly one thread calls epi for every row of tiles [t0, t1); all threads of the CTA must call tile_stream.
//
// Iteration i of the pipeline:  products of tile i (gathers issued two iterations ago) -> __syncthreads ->
// TMA issue of tile i+3 -> gathers of tile i+2 go in flight -> row sums of tile i.
template <bool SPLIT, class Epi>
__device__ __forceinline__ void tile_stream(const TileMat& M, int t0, int t1, const double* __restrict__ vec, TileCtx& cx,
                                            Epi&& epi) {
  static_assert(TS_STAGES >= 4 && 2 * TS_STAGES <= TS_RING, "the pipeline keeps three tiles resident and one in flight");
  const int tid = threadIdx.x;
  const int T = t1 - t0;
  if (T <= 0) return;
  const int stage0 = cx.stage;
  auto stage_of = [&](int i) { return (stage0 + i) % TS_STAGES; };   // int modulo by a constant

  auto issue = [&](int i) {   // thread 0; descriptor i is in the ring
    const int4 td = cx.s_td[i & (TS_RING - 1)];
    const int a4 = td.z & ~3;
    const uint32_t cnt = (uint32_t)(((td.w - a4) + 3) & ~3);
    const int s = stage_of(i);
    mbar_expect_tx(&cx.full[s], cnt * 12u);
    if (cnt) {
      bulk_g2s_hint(cx.sval + (size_t)s * TS_CAP, M.val + a4, cnt * 8u, &cx.full[s], cx.policy);
      bulk_g2s_hint(cx.sidx + (size_t)s * TS_CAP, M.idx + a4, cnt * 4u, &cx.full[s], cx.policy);
    }
  };
  // wait for tile i to land, then put its gathers in flight
  auto gather = [&](int i, double (&g)[TS_J]) {
    const int s = stage_of(i);
    mbar_wait(&cx.full[s], (cx.parity >> s) & 1u);
    cx.parity ^= 1u << s;
    const int4 td = cx.s_td[i & (TS_RING - 1)];
    const int* si = cx.sidx + (size_t)s * TS_CAP + (td.z & 3);
    const int cnt = td.w - td.z;
#pragma unroll
    for (int j = 0; j < TS_J; ++j) {
      const int kk = tid + j * NT;
      g[j] = (kk < cnt) ? ld_vec(vec + si[kk]) : 0.0;
    }
  };

  if (tid < TS_STAGES && tid < T) cx.s_td[tid] = __ldg(M.tiles + t0 + tid);
  __syncthreads();
  if (tid == 0) {
    fence_proxy_async();
    for (int i = 0; i < TS_STAGES - 1 && i < T; ++i) issue(i);
  }
  double gA[TS_J], gB[TS_J];
  gather(0, gA);
  if (T > 1) gather(1, gB);
  double carry0 = 0.0, carry1 = 0.0;   // thread 0: partial sums of a long row across its chunk tiles

  auto body = [&](int i, double (&g)[TS_J]) {
    const int s = stage_of(i);
    double* sv = cx.sval + (size_t)s * TS_CAP;
    const int4 td = cx.s_td[i & (TS_RING - 1)];
    const int k0 = td.z, a4 = k0 & ~3, off = k0 - a4, cnt = td.w - k0;
    const bool whole = td.y >= 0;
    const int R = whole ? td.y - td.x : 1;
    // rows of this tile: lanes per row (power of two), and the pointers of my first row
    int lgs = 0;
    while (lgs < 5 && ((R << lgs) << 1) <= NT) ++lgs;
    const int lg = 1 << lgs, lane = tid & (lg - 1), grp = tid >> lgs, ngrp = NT >> lgs;
    int pa = 0, pb = 0, ps = 0;
    if (whole && grp < R) {
      pa = __ldg(M.ptr + td.x + grp); pb = __ldg(M.ptr + td.x + grp + 1);
      if (SPLIT) ps = __ldg(M.split + td.x + grp);
    }
    // (A) products of tile i
#pragma unroll
    for (int j = 0; j < TS_J; ++j) {
      const int kk = tid + j * NT;
      if (kk < cnt) sv[off + kk] *= g[j];
    }
    __syncthreads();   // products visible; every thread is done with tile i-1
    if (tid == 0 && i + TS_STAGES - 1 < T) { fence_proxy_async(); issue(i + TS_STAGES - 1); }
    if (tid == 32 && i + TS_STAGES < T) cx.s_td[(i + TS_STAGES) & (TS_RING - 1)] = __ldg(M.tiles + t0 + i + TS_STAGES);
    // (B) gathers of tile i+2 go in flight
    if (i + 2 < T) gather(i + 2, g);
    // (C) row sums of tile i out of shared memory
    if (whole) {
      for (int rb = 0; rb < R; rb += ngrp) {
        const int rr = rb + grp;
        const bool valid = rr < R;
        if (rb > 0 && valid) {
          pa = __ldg(M.ptr + td.x + rr); pb = __ldg(M.ptr + td.x + rr + 1);
          if (SPLIT) ps = __ldg(M.split + td.x + rr);
        }
        double acc0 = 0.0, acc1 = 0.0;
        if (valid) {
          if (SPLIT) {
            for (int k = pa - a4 + lane; k < pb - a4; k += lg) {
              const double v = sv[k];
              if (k < ps - a4) acc0 += v; else acc1 += v;
            }
          } else {
            for (int k = pa - a4 + lane; k < pb - a4; k += lg) acc0 += sv[k];
          }
        }
        for (int o = lg >> 1; o > 0; o >>= 1) {
          acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
          if (SPLIT) acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
        }
        if (valid && lane == 0) epi(td.x + rr, acc0, acc1);
      }
    } else {
      // chunk of one long row: CTA-wide fixed-order sum, carried by thread 0 until the last chunk
      double acc0 = 0.0, acc1 = 0.0;
      const int sp = SPLIT ? __ldg(M.split + td.x) - k0 : 0;
#pragma unroll
      for (int j = 0; j < TS_J; ++j) {
        const int kk = tid + j * NT;
        if (kk < cnt) {
          const double v = sv[off + kk];
          if (SPLIT && kk >= sp) acc1 += v; else acc0 += v;
        }
      }
      acc0 = warp_combine(acc0, false);
      if (SPLIT) acc1 = warp_combine(acc1, false);
      if ((tid & 31) == 0) { cx.s_part[tid >> 5] = acc0; if (SPLIT) cx.s_part[NW + (tid >> 5)] = acc1; }
      __syncthreads();
      if (tid == 0) {
        for (int w = 0; w < NW; ++w) { carry0 += cx.s_part[w]; if (SPLIT) carry1 += cx.s_part[NW + w]; }
        if (td.y == -2) { epi(td.x, carry0, carry1); carry0 = 0.0; carry1 = 0.0; }
      }
    }
  };
  for (int i = 0; i < T; i += 2) {
    body(i, gA);
    if (i + 1 < T) body(i + 1, gB);
  }
  cx.stage = stage_of(T);
  __syncthreads();   // stages and the descriptor ring may be refilled by the next call
}

// ---- standalone kernel: y = M x on the persistent grid (C ABI b200qp_spmv_*, variant "tile") -------------
__global__ void __launch_bounds__(NT, 2)
k_spmv_tile(const TileMat M, const int* __restrict__ cta_tile, const double* __restrict__ x, double* __restrict__ y) {
  extern __shared__ __align__(128) unsigned char ts_smem[];
  __shared__ TileShared sh;
  TileCtx cx;
  tile_ctx_init(cx, ts_smem, sh);
  tile_stream<false>(M, __ldg(cta_tile + blockIdx.x), __ldg(cta_tile + blockIdx.x + 1), x, cx,
                     [&](int row, double s, double) { y[row] = s; });
}

}  // namespace b200qp
