// b200qp_csl.cu -- translation unit of the compacted-slice / TMA-ring kernels (compiled side by side with b200qp.cu).
#include "admm_kernel_csl.cuh"

namespace b200qp {

// Standalone y = M x with the production tile scheme (C ABI b200qp_spmv_*, variant "tile-tma"): the same tile blobs,
// the same producer / consumer ring and the same part_dot as phases B and Z of admm_kernel_csl -- one phase, no prefetch
// of a following stream.  Used by bench.py's per-iteration SpMV figures and for ncu captures of the SpMV alone.
__global__ void __launch_bounds__(CSL_NT, 2)
k_spmv_csl(const CslStream M, const int* __restrict__ part, int S, int capb, const double* __restrict__ x, double* __restrict__ y) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ alignas(8) uint64_t bar_full[32];
  __shared__ alignas(8) uint64_t bar_empty[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_fence_init();
    for (int s = 0; s < S; ++s) csl_mbar_arrive(&bar_empty[s]);
  }
  __syncthreads();
  const int t0 = part[blockIdx.x], T = part[blockIdx.x + 1] - t0;
  if (warp == CSL_NWC) {
    const uint64_t pol = policy_evict_first();
    int q = 0;
    for (int ib = 0; ib < T; ib += 31) {
      const unsigned mine = __ldg(M.toff + t0 + min(ib + lane, T));
      const int nt = min(31, T - ib);
      for (int i = 0; i < nt; ++i) {
        const unsigned o0 = __shfl_sync(0xffffffffu, mine, i), o1 = __shfl_sync(0xffffffffu, mine, i + 1);
        if (lane == 0) {
          const int stg = q >= S ? q - S : q;
          mbar_wait(&bar_empty[stg], q >= S ? 1u : 0u);
          const uint32_t bytes = (o1 - o0) * 16u;
          mbar_expect_tx(&bar_full[stg], bytes);
          bulk_g2s_hint(ring + (size_t)stg * capb, M.blob + (size_t)o0 * 16u, bytes, &bar_full[stg], pol);
        }
        q = q + 1 == 2 * S ? 0 : q + 1;
        __syncwarp();
      }
    }
    return;
  }
  int q = warp;
  for (int i = warp; i < T; i += CSL_NWC) {
    const int stg = q >= S ? q - S : q;
    mbar_wait(&bar_empty[stg], q >= S ? 1u : 0u);      // see admm_kernel_csl.cuh: pins full[stg] to the phase waited for
    mbar_wait(&bar_full[stg], q >= S ? 1u : 0u);
    const TileView tv = tile_view<true>(ring + (size_t)stg * capb, lane);
    double acc, dummy;
    part_dot<1, true>(tv.i1, tv.v1, tv.len1, tv.max1, x, nullptr, acc, dummy);
    __syncwarp();
    if (lane == 0) csl_mbar_arrive(&bar_empty[stg]);
    const int row = (t0 + i) * 32 + lane;
    if (row < M.rows) y[row] = acc;
    q += CSL_NWC;
    while (q >= 2 * S) q -= 2 * S;
  }
}

}  // namespace b200qp

const void* b200qp_csl_kernel(bool dist) {
  return dist ? (const void*)b200qp::admm_kernel_csl<true> : (const void*)b200qp::admm_kernel_csl<false>;
}

void b200qp_spmv_csl_launch(const b200qp::CslStream& M, const int* part, int G, int S, int capb, size_t smem, const double* x,
                            double* y, cudaStream_t stream) {
  b200qp::k_spmv_csl<<<G, b200qp::CSL_NT, smem, stream>>>(M, part, S, capb, x, y);
}
int b200qp_spmv_csl_prepare(size_t budget, size_t smem) {
  if (cudaFuncSetAttribute(b200qp::k_spmv_csl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess) return 0;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, b200qp::k_spmv_csl, b200qp::CSL_NT, smem) != cudaSuccess) return 0;
  return occ;
}
