// device_common.cuh -- shared device-side building blocks of the B200 OSQP engine.
//   * CSR matrix view, cache-hinted loads
//   * deterministic block / grid reductions (fixed order, no floating-point atomics)
//   * a watchdog-guarded grid barrier for the persistent ADMM kernel
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200qp {

constexpr int NT = 512;        // threads per CTA of the persistent kernel
constexpr int NW = NT / 32;    // warps per CTA
constexpr int KMAX = 16;       // max reduction slots per phase
constexpr double QP_INFTY = 1e30;
constexpr double INF_THRESH = 1e26;   // OSQP_INFTY * MIN_SCALING: bound counts as infinite
constexpr double RHO_MIN = 1e-6, RHO_MAX = 1e6, RHO_TOL = 1e-4;
constexpr double MIN_SCALING = 1e-4, MAX_SCALING = 1e4;

struct CsrMat {
  int rows, cols, nnz;
  const int* __restrict__ ptr;   // rows+1
  const int* __restrict__ idx;   // nnz (padded by 8 entries)
  const double* __restrict__ val;
};

// ---- loads --------------------------------------------------------------------------------
// Matrix streams are immutable while a kernel runs: read-only path, do not pollute L1.
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));   // not volatile: freely scheduled
  return v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
// Vectors change between phases of the persistent kernel and are written by other SMs:
// always fetch from L2 (L1 is not coherent across SMs).
__device__ __forceinline__ double ld_vec(const double* p) { return __ldcg(p); }

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- mbarrier / bulk-copy (TMA) PTX wrappers (sm_90+; SASS: UBLKCP, SYNCS) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Wait for a phase of an mbarrier.  One asm block with no outputs, uniform branches: the warp polls as one (the phase
// test is a warp-level SYNCS instruction on a uniform address, every lane gets the same predicate) and the compiler
// schedules around the block like around a single instruction.  Measured on the tile kernel: plain `bra` instead of
// `bra.uni` costs 3 %, a C++ loop (or any C++ branch on an output of the block) around try_wait 6-9 %.  Watchdog: a wait
// that lasts 5 s can only be a ring out of step; it ends in `trap` -- a launch failure on the host -- instead of a hung
// GPU.  NOTE for callers: the wait sees ONE parity bit, so a waiter may be at most one phase ahead of the barrier (see
// the consumer loop of admm_kernel_csl.cuh for how the tile ring guarantees that).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .u32 cnt;\n\t.reg .u64 t0, t1, dt;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "mov.u32 cnt, 0;\n\t"
      "mov.u64 t0, 0;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "add.u32 cnt, cnt, 1;\n\t"
      "setp.lt.u32 q, cnt, 256;\n\t"
      "@q bra.uni WAIT_LOOP;\n\t"
      "mov.u32 cnt, 0;\n\t"
      "mov.u64 t1, %%globaltimer;\n\t"
      "setp.eq.u64 q, t0, 0;\n\t"
      "@q mov.u64 t0, t1;\n\t"
      "sub.u64 dt, t1, t0;\n\t"
      "setp.gt.u64 q, dt, 5000000000;\n\t"
      "@q trap;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// same, with an L2 cache policy (matrix streams: evict_first, so that the gathered vectors stay L2-resident)
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// order earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// NaN-propagating max (fmax drops NaNs; the NON_CVX test needs them)
__device__ __forceinline__ double nanmax(double a, double b) {
  return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
}
__device__ __forceinline__ double combine(double a, double b, bool is_max) { return is_max ? nanmax(a, b) : a + b; }

// ---- grid-wide control block ------------------------------------------------------------------
struct Ctrl {
  unsigned long long bar;   // arrivals since launch (monotone)
  int abort;                // set by the barrier watchdog
  int pad;
};

struct GridSync {
  Ctrl* ctrl;
  int G;
  unsigned long long epoch;  // barriers passed so far (identical in every thread)
  int* s_ok;                 // shared-memory flag

  // Returns false if the watchdog fired (some CTA never arrived): every CTA then leaves the kernel.
  __device__ __forceinline__ bool barrier() {
    __syncthreads();
    epoch++;
    if (G == 1) return true;
    if (threadIdx.x == 0) {
      int ok = 1;
      __threadfence();
      atomicAdd(&ctrl->bar, 1ULL);
      const unsigned long long target = epoch * (unsigned long long)G;
      unsigned spins = 0;
      unsigned long long t0 = 0;
      while (ld_acquire_u64(&ctrl->bar) < target) {
        if ((++spins & 0xfffu) == 0) {
          if (*(volatile int*)&ctrl->abort) { ok = 0; break; }
          const unsigned long long now = global_timer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 10000000000ULL) { atomicExch(&ctrl->abort, 1); ok = 0; break; }  // 10 s
        }
      }
      __threadfence();
      *s_ok = ok;
    }
    __syncthreads();
    return *s_ok != 0;
  }
};

// ---- reductions ------------------------------------------------------------------------------
// Butterfly over the warp: every lane ends with the same value, fixed combination tree.
__device__ __forceinline__ double warp_combine(double v, bool is_max) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = combine(v, __shfl_xor_sync(0xffffffffu, v, off), is_max);
  return v;
}
template <int LG>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int off = LG / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// CTA-level: thread partials v[K] -> one value per slot, written to bank[slot*G + cta].
template <int K>
__device__ __forceinline__ void block_reduce_store(double (&v)[K], unsigned maxmask, double* bank, int G,
                                                   double* s_red /* [KMAX*NW] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double r = warp_combine(v[k], (maxmask >> k) & 1u);
    if (lane == 0) s_red[k * NW + warp] = r;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    const bool mx = (maxmask >> threadIdx.x) & 1u;
    double acc = s_red[threadIdx.x * NW];
    for (int w = 1; w < NW; ++w) acc = combine(acc, s_red[threadIdx.x * NW + w], mx);
    bank[threadIdx.x * G + blockIdx.x] = acc;
  }
}

// After the grid barrier: every CTA folds all G partials of every slot in the same order.
template <int K>
__device__ __forceinline__ void grid_reduce_load(const double* bank, int G, unsigned maxmask,
                                                 double* s_tot /* [KMAX] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = warp; k < K; k += NW) {
    const bool mx = (maxmask >> k) & 1u;
    double acc = mx ? __longlong_as_double(0xfff0000000000000LL) /* -inf */ : 0.0;
    for (int c = lane; c < G; c += 32) acc = combine(acc, __ldcg(bank + k * G + c), mx);
    acc = warp_combine(acc, mx);
    if (lane == 0) s_tot[k] = acc;
  }
  __syncthreads();
}

}  // namespace b200qp
