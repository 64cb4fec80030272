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
constexpr size_t CSL_RING_BUDGET_DEEP = 108 * 1024;   // deeper ring, 228 KB carve-out (setup measures which one wins)

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

// Multi-GPU extension of the CSL kernel (admm_kernel_csl<true>): ONE QP partitioned over `world` GPUs by tile ranges
// (rank g streams its own slab of m-side tiles and slice of n-side tiles; every vector is full length on every rank).
// The five gathered vectors p, d, w, w2, y live in peer-mapped memory: the thread that produces an element stores it
// into every peer's copy (st.relaxed.sys over NVLink / NVSwitch).  Barriers are flat: every CTA of every rank adds 1 to
// every rank's arrival counter (red.release.sys) and spins on its own GPU's counter -- one NVLink hop, no funnel through
// a leader CTA; reductions travel as per-CTA partials written into every rank's bank and are folded by every CTA in the
// same order, so all CTAs of all ranks hold identical bits and take identical decisions.
constexpr int CSL_MAXW = 8;
constexpr int CSL_GMAX = 512;               // CTAs per rank the peer-mapped banks are sized for
struct XCsl {
  int rank, world;
  int G, Gtot, goff;                         // my CTAs, CTAs of all ranks, my first slot in the banks
  const int *part_un, *part_um;              // uniform CTA partitions of [0, n) (replicated x) and [0, m) (rho re-typing)
  unsigned long long epoch0;                 // cross-GPU barriers completed by earlier launches (counters are never reset)
  unsigned long long* xflag[CSL_MAXW];       // rank h's arrival flags, one per CTA of the job [Gtot]
  double* xbank[CSL_MAXW];                   // rank h's reduction banks [2][KMAX][Gtot]
  unsigned char* xslot[CSL_MAXW];            // rank h's {value, barrier number} slots [2 sites][2][Gtot] x 16 B (fence-free barriers)
  double *p[CSL_MAXW], *d[CSL_MAXW], *w[CSL_MAXW], *w2[CSL_MAXW], *y[CSL_MAXW];
  int* abort_flag[CSL_MAXW];
  // t = rho .* (A u) as a TAGGED vector: 16-byte {value, phase tag} pairs written by the producing thread into every
  // rank's copy; the K u phase gathers it with a tag check instead of waiting at a fenced barrier after the A u phase
  int tag_t;
  unsigned char* tt[CSL_MAXW];               // rank h's tagged copy of t [m] x 16 B
  // the same for u = M^-1 r (gathered by the A u and P u products): no fenced barrier after the phases that produce it
  int tag_u;
  unsigned char* tu[CSL_MAXW];               // rank h's tagged copy of u [n] x 16 B
};

}  // namespace b200qp

// the kernel lives in its own translation unit (b200qp_csl.cu); the host driver launches it through this handle
// (dist = the multi-GPU instantiation, third kernel parameter XCsl; the single-GPU one ignores it)
const void* b200qp_csl_kernel(bool dist = false);
// standalone tile SpMV (same translation unit): occupancy query (0 = cannot run) and launch
int b200qp_spmv_csl_prepare(size_t budget, size_t smem);
void b200qp_spmv_csl_launch(const b200qp::CslStream& M, const int* part, int G, int S, int capb, size_t smem, const double* x,
                            double* y, cudaStream_t stream);
