// spmv_kernels.cuh -- standalone fp64 CSR SpMV kernels (y = M x), the hot operator of the path.
// Reference counterpart: the three SpMVs per K-apply of the indirect linear system
// (upstream algebra/cuda calls cusparseSpMV, not in tree).
//
// Variants (b200qp_spmv_variant_name):
//   0 "group"   : the in-kernel scheme of admm_kernel.cuh: LG lanes per row chosen from the mean row length, two rows
//                 per lane group, the first 3 entries per lane fetched in straight-line code (dot_pair)
//   1 "stream"  : tile_spmv.cuh -- persistent CTAs, tiles of nonzeros staged into shared memory with TMA
//                 bulk copies (cp.async.bulk + mbarrier complete_tx), nonzero-parallel gathers two
//                 tiles ahead, rows summed out of shared memory
//   2 "scalar"  : one thread per row (baseline for the roofline table)
// Algorithmic bytes per launch: 12 nnz + 4 (rows+1) + 8 rows + 8 cols  (BASELINE.md section 2).
#pragma once
#include "admm_kernel.cuh"
#include "device_common.cuh"
#include "tile_spmv.cuh"

namespace b200qp {

// ---------------------------------------------------------------------------------------- group
template <int LG>
__global__ void __launch_bounds__(256) k_spmv_group(CsrMat M, const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & (LG - 1);
  const int grp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) / LG);
  const int rowA = 2 * grp, rowB = rowA + 1;
  double accA, accB;
  dot_pair<LG, 3, 3>(M, x, rowA, rowA < M.rows, M, x, rowB, rowB < M.rows, lane, accA, accB);
  accA = group_sum<LG>(accA);
  accB = group_sum<LG>(accB);
  if (lane == 0) {
    if (rowA < M.rows) y[rowA] = accA;
    if (rowB < M.rows) y[rowB] = accB;
  }
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
