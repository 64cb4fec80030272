// spmv_kernels.cuh -- standalone fp64 CSR SpMV kernels (y = M x), the hot operator of the path.
// Reference counterpart: the three SpMVs per K-apply of the indirect linear system
// (upstream algebra/cuda calls cusparseSpMV, not in tree).
//
// Variants (b200qp_spmv_variant_name):
//   0 "group"   : LG lanes per row chosen from the mean row length, direct global loads
//   1 "stream"  : tile_spmv.cuh -- persistent CTAs, tiles of nonzeros staged into shared memory with TMA
//                 bulk copies (cp.async.bulk + mbarrier complete_tx), nonzero-parallel gathers two
//                 tiles ahead, rows summed out of shared memory
//   2 "scalar"  : one thread per row (baseline for the roofline table)
// Algorithmic bytes per launch: 12 nnz + 4 (rows+1) + 8 rows + 8 cols  (BASELINE.md section 2).
#pragma once
#include "device_common.cuh"
#include "tile_spmv.cuh"

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

}  // namespace b200qp
