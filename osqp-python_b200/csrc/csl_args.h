// csl_args.h -- launch geometry and argument block of the compacted-slice ADMM kernel (admm_kernel_csl.cuh),
// shared by its translation unit (b200qp_csl.cu) and the host driver (b200qp.cu).
#pragma once
#include "admm_kernel.cuh"

namespace b200qp {

constexpr int CSL_NWC = 11;                 // consumer warps per CTA
constexpr int CSL_NT = CSL_NWC * 32 + 32;  // + one producer warp
constexpr int CSL_NW = CSL_NT / 32;
constexpr int CSL_CT = CSL_NWC * 32;       // consumer threads
constexpr int CSL_HDR = 80;                // header + two length arrays
constexpr size_t CSL_RING_BUDGET = 94 * 1024;   // dynamic shared memory per CTA (two CTAs per SM inside a 196 KB carve-out)

struct CslStream {
  int rows, ntiles;
  const unsigned* toff;          // ntiles + 1 offsets into blob, in 16-byte units
  const unsigned char* blob;
};

struct CArgs {                   // what the CSL kernel needs beyond KArgs
  CslStream N, M;                // n-side tiles (P | A'), m-side tiles (A)
  const int *part_tn, *part_tm;  // G + 1 tile boundaries per stream
  double *ps, *ss;               // n-vectors: PCG search direction and K * direction
  int S, capb;                   // ring: stages and bytes per stage
  int direct;                    // diagnostics: 1 = no TMA ring, plain loads of the tiles
  long long* cta_ns;             // optional [G][2]: per-CTA time inside the B / C phases before the barrier (diagnostics)
};

}  // namespace b200qp

// the kernel lives in its own translation unit (b200qp_csl.cu); the host driver launches it through this handle
const void* b200qp_csl_kernel();
// standalone tile SpMV (same translation unit): occupancy query (0 = cannot run) and launch
int b200qp_spmv_csl_prepare(size_t budget, size_t smem);
void b200qp_spmv_csl_launch(const b200qp::CslStream& M, const int* part, int G, int S, int capb, size_t smem, const double* x,
                            double* y, cudaStream_t stream);
