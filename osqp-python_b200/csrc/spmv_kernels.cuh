// spmv_kernels.cuh -- standalone fp64 CSR SpMV kernels (y = M x), the hot operator of the path.
// Reference counterpart: the three SpMVs per K-apply of the indirect linear system
// (upstream algebra/cuda calls cusparseSpMV, not in tree).
//
// Variants (b200qp_spmv_variant_name):
//   0 "group"   : LG lanes per row chosen from the mean row length, direct global loads
//   1 "stream"  : persistent CTAs, row tiles staged into shared memory with TMA bulk copies
//                 (cp.async.bulk + mbarrier complete_tx), multi-stage ring; rows are then
//                 reduced out of shared memory while x is gathered from L2
//   2 "scalar"  : one thread per row (baseline for the roofline table)
// Algorithmic bytes per launch: 12 nnz + 4 (rows+1) + 8 rows + 8 cols  (BASELINE.md section 2).
#pragma once
#include "device_common.cuh"

namespace b200qp {

// ---------------------------------------------------------------------------------------- group
template <int LG>
__global__ void __launch_bounds__(256) k_spmv_group(CsrMat M, const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & (LG - 1);
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) / LG);
  double acc = 0.0;
  if (row < M.rows) {
    const int a = __ldg(M.ptr + row), b = __ldg(M.ptr + row + 1);
#pragma unroll 4
    for (int k = a + lane; k < b; k += LG) acc = fma(ld_stream(M.val + k), ld_vec(x + ld_stream(M.idx + k)), acc);
  }
  acc = group_sum<LG>(acc);
  if (row < M.rows && lane == 0) y[row] = acc;
}

__global__ void __launch_bounds__(256) k_spmv_scalar(CsrMat M, const double* __restrict__ x, double* __restrict__ y) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row < M.rows) {
    const int a = __ldg(M.ptr + row), b = __ldg(M.ptr + row + 1);
    double acc = 0.0;
    for (int k = a; k < b; ++k) acc = fma(ld_stream(M.val + k), ld_vec(x + ld_stream(M.idx + k)), acc);
    y[row] = acc;
  }
}

// --------------------------------------------------------------------------------------- stream
// mbarrier / bulk-copy PTX wrappers (sm_90+; SASS: UBLKCP, SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int ST_NT = 512;        // threads per CTA
constexpr int ST_TILE = 2048;     // max nnz per staged tile
constexpr int ST_CAP = ST_TILE + 8;
constexpr int ST_STAGES = 4;
constexpr size_t ST_SMEM = (size_t)ST_STAGES * ST_CAP * 12;

// tile_row[t] .. tile_row[t+1] is the row range of tile t (built by make_tiles on the host):
// at most ST_NT rows and ST_TILE nnz, except single rows longer than ST_TILE.
__global__ void __launch_bounds__(ST_NT, 2)
k_spmv_stream(CsrMat M, const int* __restrict__ tile_row, int T, const double* __restrict__ x, double* __restrict__ y) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sval = reinterpret_cast<double*>(smem_raw);
  int* sidx = reinterpret_cast<int*>(smem_raw + (size_t)ST_STAGES * ST_CAP * 8);
  __shared__ __align__(8) uint64_t full[ST_STAGES];
  __shared__ double s_part[ST_NT / 32];

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < ST_STAGES; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int G = gridDim.x;
  const int my = (blockIdx.x < T) ? (T - blockIdx.x + G - 1) / G : 0;

  auto issue = [&](int i) {
    const int t = blockIdx.x + i * G, s = i % ST_STAGES;
    const int a = __ldg(M.ptr + __ldg(tile_row + t)), b = __ldg(M.ptr + __ldg(tile_row + t + 1));
    if (b - a > ST_TILE) return;                 // long row: read directly, nothing staged
    const int a4 = a & ~3;
    const uint32_t cnt = (uint32_t)(((b - a4) + 3) & ~3);
    mbar_expect_tx(&full[s], cnt * 12u);
    if (cnt) {
      bulk_g2s(sval + (size_t)s * ST_CAP, M.val + a4, cnt * 8u, &full[s]);
      bulk_g2s(sidx + (size_t)s * ST_CAP, M.idx + a4, cnt * 4u, &full[s]);
    }
  };
  if (tid == 0)
    for (int i = 0; i < ST_STAGES - 1 && i < my; ++i) issue(i);

  uint32_t phase_bits = 0;   // parity per stage
  for (int i = 0; i < my; ++i) {
    if (tid == 0 && i + ST_STAGES - 1 < my) issue(i + ST_STAGES - 1);
    const int t = blockIdx.x + i * G, s = i % ST_STAGES;
    const int r0 = __ldg(tile_row + t), r1 = __ldg(tile_row + t + 1);
    const int a = __ldg(M.ptr + r0), b = __ldg(M.ptr + r1);
    if (b - a > ST_TILE) {
      // single long row: whole CTA streams it from global memory
      double acc = 0.0;
      for (int k = a + tid; k < b; k += ST_NT) acc = fma(ld_stream(M.val + k), ld_vec(x + ld_stream(M.idx + k)), acc);
      acc = warp_combine(acc, false);
      if ((tid & 31) == 0) s_part[tid >> 5] = acc;
      __syncthreads();
      if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < ST_NT / 32; ++w) tot += s_part[w];
        y[r0] = tot;
      }
    } else {
      mbar_wait(&full[s], (phase_bits >> s) & 1u);
      phase_bits ^= (1u << s);
      const int a4 = a & ~3;
      const double* sv = sval + (size_t)s * ST_CAP;
      const int* si = sidx + (size_t)s * ST_CAP;
      const int R = r1 - r0;
      int lg = 1;
      while (lg < 32 && R * lg * 2 <= ST_NT) lg <<= 1;       // lanes per row for this tile
      const int lane = tid & (lg - 1), grp = tid / lg, ngrp = ST_NT / lg;
      for (int rb = 0; rb < R; rb += ngrp) {
        const int rr = rb + grp;
        double acc = 0.0;
        if (rr < R) {
          const int ka = __ldg(M.ptr + r0 + rr) - a4, kb = __ldg(M.ptr + r0 + rr + 1) - a4;
#pragma unroll 4
          for (int k = ka + lane; k < kb; k += lg) acc = fma(sv[k], ld_vec(x + si[k]), acc);
        }
        for (int off = lg >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (rr < R && lane == 0) y[r0 + rr] = acc;
      }
    }
    __syncthreads();   // stage s may be refilled by the next issue()
  }
}

}  // namespace b200qp
