// admm_kernel_csl.cuh -- round-2 persistent ADMM kernel: compacted-slice matrix tiles streamed by TMA.
//
// Same algorithm, settings, result block and device vectors as admm_kernel.cuh (reference loop:
// src/osqppurepy/_osqp.py:1188-1306); what changes is how the three sparse matrices reach the SMs.
//
// Layout (built once at setup, csl_host.h): rows are cut into tiles of 32 -- one warp, lane t owns row t.  Inside a
// tile the entries are stored column-major and compacted (entry 0 of every row that has one, in row order, then
// entry 1, ...): the loads of one "column" are contiguous for the warp, nothing is padded, no row pointer is
// chased and no shuffle reduction is needed; a row's dot product is summed by its own lane in a fixed order.
// An n-side tile carries the 32 rows of P *and* of A' (K p = P p + sigma p + A' w walks them together), an m-side
// tile the 32 rows of A.  Tile = one contiguous, 16-byte aligned blob:
//     int4 {cnt1, cnt2, maxlen1, maxlen2} | u8 len1[32] | u8 len2[32] | f64 val1[cnt1] | f64 val2[cnt2] | i32 idx1[cnt1] | i32 idx2[cnt2]
// so ONE cp.async.bulk (TMA, SASS UBLKCP) brings a tile into shared memory.
//
// Execution: 2 CTAs per SM, each 8 consumer warps + 1 producer warp.  The producer lane streams the CTA's tiles
// into a ring of shared-memory stages (mbarrier complete_tx "full" / consumer-arrive "empty"); consumer warp w takes
// tiles w, w+8, ... with no CTA-wide barrier inside a phase.  The matrix streams of consecutive phases strictly
// alternate (A, then P|A', then A, ...), so when the producer has issued the last tile of a phase it goes straight
// on with the first tiles of the next one: they land while the CTAs sit in the grid barrier.  Phases that break the
// alternation (the rho-update refresh, the infeasibility-certificate norms) read their tiles with plain loads.
// Gathers (ld.global.cg from the L2-resident vectors) are issued U at a time per lane before the first FMA.
//
// PCG: single-reduction recurrence (Chronopoulos & Gear), three grid barriers per iteration:
//   B  m-tiles: t = rho .* (A u)         C  n-tiles: w = P u + sigma u + A' t, <w,u>
//   V  n-elements: p, s, d, r, u = M^-1 r, <r,u>, |r|inf       (128-bit loads / stores)
#pragma once
#include <type_traits>

#include "admm_kernel.cuh"
#include "csl_args.h"

namespace b200qp {

__device__ __forceinline__ void csl_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// predicated gather: inactive lanes issue no memory request (a shared dummy address turns one L2 slice into a hot spot)
__device__ __forceinline__ double csl_gather_if(const double* p, bool act) {
  double v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@p ld.global.cg.f64 %0, [%1];\n\t}"
               : "=d"(v) : "l"(p), "r"((int)act));
  return v;
}

// pointers into one tile (shared memory stage or global blob)
struct TileView {
  const double *v1, *v2;
  const int *i1, *i2;
  int len1, len2, max1, max2;
};
template <bool SMEM>
__device__ __forceinline__ TileView tile_view(const unsigned char* tp, int lane) {
  TileView t;
  int4 h;
  if (SMEM) h = *reinterpret_cast<const int4*>(tp);
  else h = __ldg(reinterpret_cast<const int4*>(tp));
  t.max1 = h.z; t.max2 = h.w;
  t.len1 = SMEM ? tp[16 + lane] : __ldg(tp + 16 + lane);
  t.len2 = SMEM ? tp[48 + lane] : __ldg(tp + 48 + lane);
  t.v1 = reinterpret_cast<const double*>(tp + CSL_HDR);
  t.v2 = t.v1 + h.x;
  t.i1 = reinterpret_cast<const int*>(t.v2 + h.y);
  t.i2 = t.i1 + h.x;
  return t;
}

// One part of a tile: lane-private row dot products against NV vectors, five "columns" at a time: slot numbers
// (ballot-compaction; none needed when every row of the tile has maxlen entries), all index loads, all gathers, all
// value loads, then the FMAs in DESCENDING order -- the first FMA needs the last gather, which keeps ptxas from
// sinking every load next to its use.  (Ten at a time was measured: same rate, 1 KB of spills at 128 registers.)
template <int NV, bool SMEM, int U = 5>
__device__ __forceinline__ void part_dot(const int* ip, const double* vp, int len, int maxlen, const double* __restrict__ va,
                                         const double* __restrict__ vb, double& sa, double& sb, int j_begin = 0, int pos0 = 0) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const bool uniform = __all_sync(0xffffffffu, len == maxlen);
  int pos = pos0;
  sa = 0.0; sb = 0.0;
  for (int j0 = j_begin; j0 < maxlen; j0 += U) {
    int k[U], c[U];
    double ga[U], gb[NV > 1 ? U : 1], v[U];
    if (uniform) {
#pragma unroll
      for (int u = 0; u < U; ++u) k[u] = (len > j0 + u) ? (j0 + u) * 32 + lane : 0;
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool act = len > j0 + u;
        const unsigned mask = __ballot_sync(0xffffffffu, act);
        k[u] = act ? pos + __popc(mask & lt) : 0;     // slot 0 always exists when maxlen > 0
        pos += __popc(mask);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = SMEM ? ip[k[u]] : ld_stream(ip + k[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      ga[u] = csl_gather_if(va + c[u], len > j0 + u);
      if (NV > 1) gb[u] = csl_gather_if(vb + c[u], len > j0 + u);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = SMEM ? vp[k[u]] : ld_stream(vp + k[u]);
#pragma unroll
    for (int u = U - 1; u >= 0; --u) {
      sa = fma(v[u], ga[u], sa);
      if (NV > 1) sb = fma(v[u], gb[u], sb);
    }
  }
}
// Same dot products against a TAGGED vector (multi-GPU kernel): element i lives in a 16-byte {value, tag} pair that its
// producer -- on this GPU or on a peer, over NVLink -- writes with one aligned vector store.  A gather that finds another
// tag than the one of the current phase simply repeats: the consumer phase needs no barrier (and the producer no
// system-scope fence) in front of it, late elements cost a few re-loads of exactly those elements.
__device__ __forceinline__ void csl_ld_pair(const unsigned char* p, double& v, unsigned long long& e) {
  long long bits;
  asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(bits), "=l"(e) : "l"(p) : "memory");
  v = __longlong_as_double(bits);
}
__device__ __forceinline__ void part_dot_tag(const int* ip, const double* vp, int len, int maxlen, const unsigned char* __restrict__ tagged,
                                             unsigned long long tag, double& sa) {
  constexpr int U = 5;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const bool uniform = __all_sync(0xffffffffu, len == maxlen);
  int pos = 0;
  sa = 0.0;
  for (int j0 = 0; j0 < maxlen; j0 += U) {
    int k[U], c[U];
    double g[U], v[U];
    unsigned long long e[U];
    if (uniform) {
#pragma unroll
      for (int u = 0; u < U; ++u) k[u] = (len > j0 + u) ? (j0 + u) * 32 + lane : 0;
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool act = len > j0 + u;
        const unsigned mask = __ballot_sync(0xffffffffu, act);
        k[u] = act ? pos + __popc(mask & lt) : 0;
        pos += __popc(mask);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = ip[k[u]];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      g[u] = 0.0; e[u] = tag;
      if (len > j0 + u) csl_ld_pair(tagged + (size_t)c[u] * 16, g[u], e[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = vp[k[u]];
#pragma unroll
    for (int u = U - 1; u >= 0; --u) {
      while (e[u] != tag) csl_ld_pair(tagged + (size_t)c[u] * 16, g[u], e[u]);     // rare: the element is still on its way
      sa = fma(v[u], g[u], sa);
    }
  }
}
// n-side tile, both parts at once: returns (A' part . vec2) + sign * (P part . vec1).  The first five columns of the
// P part (usually all of it) are loaded and gathered BEFORE the A' part runs and folded in after it, chained onto
// its sum so that ptxas cannot pull those FMAs in front of the A' gathers: one latency instead of two per tile.
template <bool SMEM>
__device__ __forceinline__ double pair_dot(const TileView& tv, const double* __restrict__ vec1, const double* __restrict__ vec2,
                                           double sign) {
  const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
  int pos = 0;
  int k[5], c[5];
  double g[5];
  const bool any1 = tv.max1 > 0;
  if (any1) {
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const bool act = tv.len1 > u;
      const unsigned mask = __ballot_sync(0xffffffffu, act);
      k[u] = act ? pos + __popc(mask & lt) : 0;
      pos += __popc(mask);
    }
#pragma unroll
    for (int u = 0; u < 5; ++u) c[u] = SMEM ? tv.i1[k[u]] : ld_stream(tv.i1 + k[u]);
#pragma unroll
    for (int u = 0; u < 5; ++u) g[u] = csl_gather_if(vec1 + c[u], tv.len1 > u);
  }
  double acc, dummy;
  part_dot<1, SMEM>(tv.i2, tv.v2, tv.len2, tv.max2, vec2, nullptr, acc, dummy);
  if (any1) {
    // the P values are fetched only now: five doubles fewer to hold through the A' rounds
#pragma unroll
    for (int u = 4; u >= 0; --u) acc = fma(sign * (SMEM ? tv.v1[k[u]] : ld_stream(tv.v1 + k[u])), g[u], acc);
    if (tv.max1 > 5) {
      double rest;
      part_dot<1, SMEM>(tv.i1, tv.v1, tv.len1, tv.max1, vec1, nullptr, rest, dummy, 5, pos);
      acc = fma(sign, rest, acc);
    }
  }
  return acc;
}

// sum of val^2 * vec[idx] (Jacobi diagonal); rare path, plain loop
template <bool SMEM>
__device__ __forceinline__ double part_dot_sq(const int* ip, const double* vp, int len, int maxlen, const double* __restrict__ va) {
  const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
  int pos = 0;
  double s = 0.0;
  for (int j = 0; j < maxlen; ++j) {
    const bool act = len > j;
    const unsigned mask = __ballot_sync(0xffffffffu, act);
    if (act) {
      const int k = pos + __popc(mask & lt);
      const double v = SMEM ? vp[k] : ld_stream(vp + k);
      s = fma(v * v, __ldcg(va + (SMEM ? ip[k] : ld_stream(ip + k))), s);
    }
    pos += __popc(mask);
  }
  return s;
}

// After the barrier: every CTA folds all Gt partials of every slot with ALL its threads (thread t takes t, t + NT, ...;
// warp butterfly; warps in order) -- one L2 round trip per slot whatever Gt is (8 GPUs: 2368 partials), fixed order.
template <int K, int NWARPS>
__device__ __forceinline__ void grid_reduce_load_all(const double* bank, int Gt, unsigned maxmask, double* s_tot, double* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const bool mx = (maxmask >> k) & 1u;
    acc[k] = mx ? __longlong_as_double(0xfff0000000000000LL) : 0.0;
    for (int cc = threadIdx.x; cc < Gt; cc += NWARPS * 32) acc[k] = combine(acc[k], __ldcg(bank + (size_t)k * Gt + cc), mx);
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double r = warp_combine(acc[k], (maxmask >> k) & 1u);
    if (lane == 0) s_red[k * NWARPS + warp] = r;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    const bool mx = (maxmask >> threadIdx.x) & 1u;
    double t = s_red[threadIdx.x * NWARPS];
    for (int w = 1; w < NWARPS; ++w) t = combine(t, s_red[threadIdx.x * NWARPS + w], mx);
    s_tot[threadIdx.x] = t;
  }
  __syncthreads();
}

// system-scope operations of the multi-GPU instantiation (peer-mapped memory over NVLink / NVSwitch)
__device__ __forceinline__ void csl_st_sys(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void csl_st_sys2(double2* p, double2 v) {
  asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void csl_st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// 16-byte {value, epoch} slots: one aligned vector store / load each, so a reader that sees the epoch sees the value
__device__ __forceinline__ void csl_st_slot(void* p, double v, unsigned long long e) {
  asm volatile("st.relaxed.sys.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(e) : "memory");
}
__device__ __forceinline__ void csl_ld_slot(const void* p, double& v, unsigned long long& e) {
  long long bits;
  asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(bits), "=l"(e) : "l"(p) : "memory");
  v = __longlong_as_double(bits);
}
__device__ __forceinline__ unsigned long long csl_ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <bool DIST>
__global__ void __launch_bounds__(CSL_NT, 2)
admm_kernel_csl(const __grid_constant__ KArgs a, const __grid_constant__ CArgs c, const __grid_constant__ XCsl x) {
  extern __shared__ __align__(128) unsigned char ring[];   // [S][capb]
  __shared__ double s_red[KMAX * CSL_NW];
  __shared__ double s_tot[KMAX];
  __shared__ int s_ok;
  __shared__ alignas(8) uint64_t bar_full[32];
  __shared__ alignas(8) uint64_t bar_empty[32];

  const int G = a.G, cta = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool producer = warp == CSL_NWC;
  const int S = c.S;
  // tile ranges of this CTA; element ranges follow the tiles
  const int tn0 = c.part_tn[cta], tn1 = c.part_tn[cta + 1], TN = tn1 - tn0;
  const int tm0 = c.part_tm[cta], tm1 = c.part_tm[cta + 1], TM = tm1 - tm0;
  const int n0 = min(tn0 * 32, a.n), n1 = min(tn1 * 32, a.n);
  const int m0 = min(tm0 * 32, a.m), m1 = min(tm1 * 32, a.m);
  // multi-GPU: the tile ranges cover this rank's slab / slice only; x is replicated (every rank updates all of it) and
  // every rank re-derives the whole rho vector, over uniform shares
  const int un0 = DIST ? x.part_un[cta] : n0, un1 = DIST ? x.part_un[cta + 1] : n1;
  const int um0 = DIST ? x.part_um[cta] : m0, um1 = DIST ? x.part_um[cta + 1] : m1;
  const int Gt = DIST ? x.Gtot : G;          // partials per reduction slot
  GridSync sync{a.ctrl, G, 0ULL, &s_ok};
  const double NEG_INF = __longlong_as_double(0xfff0000000000000LL);

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_fence_init();
  }
  __syncthreads();

  const double alpha = a.alpha, sigma = a.sigma;
  const double cinv = 1.0 / a.c;
  const double c_eff = a.unscale ? a.c : 1.0, cinv_eff = a.unscale ? cinv : 1.0;
  const double* Dinv_t = a.unscale ? a.Dinv : nullptr;
  const double* D_t = a.unscale ? a.D : nullptr;
  const double* Einv_t = a.unscale ? a.Einv : nullptr;
  const double* E_t = a.unscale ? a.E : nullptr;

  double rho = a.rho;
  int rho_updates = 0, status = 0, k = 0;
  long long cg_total = 0;
  const unsigned long long t_start = global_timer_ns();

  // Residual information of the last check: identical in every thread and touched a few times per check only, so it
  // lives in shared memory (every thread stores the same value) instead of 42 registers per thread.
  __shared__ double cold[24];
  double &obj = cold[0], &dobj = cold[1], &gap = cold[2], &prim_res = cold[3], &dual_res = cold[4], &nprim = cold[5],
         &ndual = cold[6], &s_prim = cold[7], &s_nprim = cold[8], &s_dual = cold[9], &s_ndual = cold[10], &ndy = cold[11],
         &lhs = cold[12], &vu = cold[13], &vl = cold[14], &sup = cold[15], &t_prim = cold[16], &t_ax = cold[17],
         &t_z = cold[18], &ndx = cold[19], &qdx = cold[20];
  if (tid < 24) cold[tid] = (tid == 13 || tid == 14) ? NEG_INF : 0.0;
  __syncthreads();
  double rz = 0, r0max = 0;
  bool time_up = false, tl_pending = false;

  // ---- how the CTAs of the job meet: grid barrier (one GPU) or flat cross-GPU barrier; partial banks; pushes
  const size_t xbank_par = DIST ? (size_t)KMAX * (size_t)x.Gtot : 0;
  auto barrier = [&]() -> bool {
    if constexpr (!DIST) {
      return sync.barrier();
    } else {
      __syncthreads();                       // every thread's stores of this phase (pushes included) are issued
      sync.epoch++;
      const unsigned long long e = x.epoch0 + sync.epoch;
      // arrival: one flag per CTA per rank, written with a release store (cumulative: the CTA's earlier peer stores --
      // pushes and partials -- are visible before it).  No atomics: 296 remote atomics on one address serialise in the
      // peer's L2 (measured: 17-20 us per barrier; a two-level barrier with a last-arriver hand-off: 25 us), 296 stores to
      // 296 addresses pipeline over NVLink (10-12 us, most of it the system-scope fence of the release).
      if (tid < CSL_MAXW && tid < x.world) {
        unsigned long long* dst = x.xflag[0];
#pragma unroll
        for (int h = 1; h < CSL_MAXW; ++h) if (tid == h) dst = x.xflag[h];
        csl_st_release_sys(dst + x.goff + cta, e);
      }
      // wait: thread t watches flags t, t + NT, ... of MY rank's array (the peers' CTAs write there) with acquire loads.
      // (Measured alternatives: relaxed polls + one system fence per thread 2x slower -- concurrent MEMBAR.SYS are what
      // costs; warp 0 polling all flags with a back-off 5x slower; byte flags, eight per acquire load, 4-8 % slower --
      // byte-enabled remote stores.)
      const unsigned long long* mine = x.xflag[0];
#pragma unroll
      for (int h = 1; h < CSL_MAXW; ++h) if (x.rank == h) mine = x.xflag[h];
      int ok = 1;
      for (int f = tid; f < x.Gtot; f += CSL_NT) {
        unsigned spins = 0;
        unsigned long long t0 = 0;
        while (csl_ld_acquire_sys(mine + f) < e) {
          if ((++spins & 0x3ffu) == 0) {
            if (*(volatile int*)&a.ctrl->abort) { ok = 0; break; }
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 10000000000ULL) { ok = 0; break; }   // 10 s watchdog
          }
        }
        if (!ok) break;
      }
      ok = __syncthreads_and(ok);
      if (!ok && tid == 0) {     // tell everybody, so that no rank spins until its own watchdog
#pragma unroll
        for (int h = 0; h < CSL_MAXW; ++h) if (h < x.world) *(volatile int*)x.abort_flag[h] = 1;
      }
      return ok != 0;
    }
  };
  // CTA partials of K slots -> the bank every CTA will fold after the next barrier (on every rank when DIST)
  auto reduce_store = [&](auto& v, auto k_tag, unsigned maxmask) {
    constexpr int K = decltype(k_tag)::value;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double r = warp_combine(v[k], (maxmask >> k) & 1u);
      if (lane == 0) s_red[k * CSL_NW + warp] = r;
    }
    __syncthreads();
    if (tid < K) {
      const bool mx = (maxmask >> tid) & 1u;
      double acc = s_red[tid * CSL_NW];
      for (int w = 1; w < CSL_NW; ++w) acc = combine(acc, s_red[tid * CSL_NW + w], mx);
      if constexpr (DIST) {
        const size_t off = (size_t)(sync.epoch & 1ULL) * xbank_par + (size_t)tid * x.Gtot + x.goff + cta;
#pragma unroll
        for (int h = 0; h < CSL_MAXW; ++h) if (h < x.world) csl_st_sys(x.xbank[h] + off, acc);
      } else {
        a.red[(size_t)(sync.epoch & 1ULL) * KMAX * G + tid * G + cta] = acc;
      }
    }
  };
  // fold after the barrier (the bank index is the parity the partials were stored under: epoch has advanced by one)
  auto reduce_load = [&](auto k_tag, unsigned maxmask) {
    constexpr int K = decltype(k_tag)::value;
    const unsigned long long par = (sync.epoch - 1ULL) & 1ULL;
    const double* bk;
    if constexpr (DIST) {
      const double* mine = x.xbank[0];
#pragma unroll
      for (int h = 1; h < CSL_MAXW; ++h) if (x.rank == h) mine = x.xbank[h];
      bk = mine + (size_t)par * xbank_par;
    } else {
      bk = a.red + (size_t)par * KMAX * G;
    }
    grid_reduce_load_all<K, CSL_NW>(bk, Gt, maxmask, s_tot, s_red);
  };
  // A barrier whose phase pushed nothing (C: only <w,u> leaves the CTA) needs no release fence at all: every CTA writes
  // its partial as a self-validating {value, barrier number} slot into every rank's slot array and every CTA folds the
  // slots as they turn up -- one NVLink hop, no MEMBAR.SYS.  Having every partial also means every CTA of the job has
  // finished the phase (a CTA stores its partial after its last tile), which is all the next phase relies on: V only
  // touches the CTA's own elements of w, and overwrites u only after every gather of the old u has returned.
  // `site` selects the slot bank: two payload barriers that follow each other with no fenced barrier in between (C, then
  // V in tagged-u mode) must not share slots -- a fast CTA would overwrite its slot for the second barrier while a slow
  // one is still folding the first.  One bank per call site is enough: a site's slots are rewritten only after the other
  // site's barrier, which every CTA enters after it has finished folding this one.
  auto barrier_payload = [&](auto& v, auto k_tag, unsigned maxmask, int site) -> bool {
    constexpr int K = decltype(k_tag)::value;
    static_assert(K <= 2, "slot array holds two slots");
    if constexpr (!DIST) {
      reduce_store(v, k_tag, maxmask);
      if (!barrier()) return false;
      reduce_load(k_tag, maxmask);
      return true;
    } else {
      const unsigned long long e = x.epoch0 + sync.epoch + 1ULL;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const double r = warp_combine(v[k], (maxmask >> k) & 1u);
        if (lane == 0) s_red[k * CSL_NW + warp] = r;
      }
      __syncthreads();
      if (tid < K) {
        const bool mx = (maxmask >> tid) & 1u;
        double acc = s_red[tid * CSL_NW];
        for (int w = 1; w < CSL_NW; ++w) acc = combine(acc, s_red[tid * CSL_NW + w], mx);
        const size_t off = ((size_t)(site * 2 + tid) * x.Gtot + x.goff + cta) * 16;
#pragma unroll
        for (int h = 0; h < CSL_MAXW; ++h) if (h < x.world) csl_st_slot(x.xslot[h] + off, acc, e);
      }
      sync.epoch++;
      const unsigned char* mine = x.xslot[0];
#pragma unroll
      for (int h = 1; h < CSL_MAXW; ++h) if (x.rank == h) mine = x.xslot[h];
      int ok = 1;
      double acc[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const bool mx = (maxmask >> k) & 1u;
        acc[k] = mx ? NEG_INF : 0.0;
        for (int cc = tid; cc < x.Gtot && ok; cc += CSL_NT) {
          const unsigned char* sp = mine + ((size_t)(site * 2 + k) * x.Gtot + cc) * 16;
          double val; unsigned long long ep;
          unsigned spins = 0;
          unsigned long long t0 = 0;
          for (;;) {
            csl_ld_slot(sp, val, ep);
            if (ep == e) break;
            if ((++spins & 0x3ffu) == 0) {
              if (*(volatile int*)&a.ctrl->abort) { ok = 0; break; }
              const unsigned long long now = global_timer_ns();
              if (t0 == 0) t0 = now;
              else if (now - t0 > 10000000000ULL) { ok = 0; break; }
            }
          }
          if (ok) acc[k] = combine(acc[k], val, mx);
        }
      }
      __syncthreads();                       // s_red is free again (the partial stores above read it)
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const double r = warp_combine(acc[k], (maxmask >> k) & 1u);
        if (lane == 0) s_red[k * CSL_NW + warp] = r;
      }
      ok = __syncthreads_and(ok);
      if (tid < K) {
        const bool mx = (maxmask >> tid) & 1u;
        double t = s_red[tid * CSL_NW];
        for (int w = 1; w < CSL_NW; ++w) t = combine(t, s_red[tid * CSL_NW + w], mx);
        s_tot[tid] = t;
      }
      __syncthreads();
      if (!ok && tid == 0) {
#pragma unroll
        for (int h = 0; h < CSL_MAXW; ++h) if (h < x.world) *(volatile int*)x.abort_flag[h] = 1;
      }
      return ok != 0;
    }
  };
  // all-gather by stores: the producing thread writes its element into every peer's copy
#define CSL_PUSH(tab, i, v) do { if constexpr (DIST) { _Pragma("unroll") for (int h_ = 0; h_ < CSL_MAXW; ++h_) if (h_ < x.world && h_ != x.rank) csl_st_sys(x.tab[h_] + (i), (v)); } } while (0)
#define CSL_PUSH2(tab, i2, v) do { if constexpr (DIST) { _Pragma("unroll") for (int h_ = 0; h_ < CSL_MAXW; ++h_) if (h_ < x.world && h_ != x.rank) csl_st_sys2(reinterpret_cast<double2*>(x.tab[h_]) + (i2), (v)); } } while (0)

  // u = M^-1 r is gathered by the A u and P u products of every rank.  Tagged mode (multi-GPU): the producer writes
  // {u, phase tag} pairs into every rank's tagged copy (own included) and keeps the plain array for its own elements;
  // the gathers check the tag, so the phase that follows a producer of u needs neither a fence nor the pushes to have
  // landed.  tag_u = the barrier count at which the current u was produced (identical in every thread of the job).
  unsigned long long tag_u = 0;
  const bool tagged_u = DIST && x.tag_u != 0;
#define CSL_PUT_U(i, v) do { a.p[(i)] = (v); if constexpr (DIST) { if (tagged_u) { _Pragma("unroll") for (int h_ = 0; h_ < CSL_MAXW; ++h_) if (h_ < x.world) csl_st_slot(x.tu[h_] + (size_t)(i) * 16, (v), tag_u); } else { CSL_PUSH(p, (i), (v)); } } } while (0)
  const unsigned char* tu_mine = nullptr;
  if constexpr (DIST) {
    tu_mine = x.tu[0];
#pragma unroll
    for (int h = 1; h < CSL_MAXW; ++h) if (x.rank == h) tu_mine = x.tu[h];
  }

  // ---- ring state -----------------------------------------------------------------------------------------------
  // ring_q = tiles issued / consumed so far, modulo 2S (stage = q % S, parity = q >= S); identical in every thread.
  // "empty" barriers get one arrival up front, so use k of a stage waits for parity k & 1 on both barriers.
  int ring_q = 0;
  int pre_issued = 0;                 // tiles of the upcoming ring phase already issued by the producer
  const uint64_t pol = policy_evict_first();
  const int KPRE = S - CSL_NWC;       // tiles of the next stream issued ahead: their stages free up before the phase ends
  if (tid == 0) for (int s = 0; s < S; ++s) csl_mbar_arrive(&bar_empty[s]);
  __syncthreads();

  // producer warp: issue tiles [i0, i1) of a stream (first tile index t0) whose tile 0 has ring sequence number q0
  // ring position helpers (no integer division on the per-tile path): q in [0, 2S), stage = q mod S, parity = q >= S
  auto q_add = [&](int q, int d) { q += d; while (q >= 2 * S) q -= 2 * S; return q; };   // d < 2S per call on the hot paths
  auto issue = [&](const CslStream& st, int t0, int i0, int i1, int q0) {
    int q = (q0 + i0) % (2 * S);
    for (int ib = i0; ib < i1; ib += 31) {
      const unsigned mine = __ldg(st.toff + t0 + min(ib + lane, i1));   // 32 offsets = 31 tiles per round
      const int nt = min(31, i1 - ib);
      for (int i = 0; i < nt; ++i) {
        const unsigned o0 = __shfl_sync(0xffffffffu, mine, i), o1 = __shfl_sync(0xffffffffu, mine, i + 1);
        if (lane == 0) {
          const int stg = q >= S ? q - S : q;
          const unsigned par = q >= S ? 1u : 0u;
          mbar_wait(&bar_empty[stg], par);          // previous occupant released (or the initial arrival)
          const uint32_t bytes = (o1 - o0) * 16u;
          mbar_expect_tx(&bar_full[stg], bytes);
          bulk_g2s_hint(ring + (size_t)stg * c.capb, st.blob + (size_t)o0 * 16u, bytes, &bar_full[stg], pol);
        }
        q = q + 1 == 2 * S ? 0 : q + 1;
        __syncwarp();
      }
    }
  };
  // One ring phase over stream `cur` (T tiles from tile t0; the first pre_issued already in flight); then the
  // producer goes on with the first tiles of `nxt`.  body(tile, view, done, smem_tag): called by all lanes of the
  // consumer warp that owns the tile; done() hands the stage back, once, after the last read of the view.
  auto ring_phase = [&](const CslStream& cur, int t0, int T, const CslStream& nxt, int nt0, int Tnext, auto&& body) {
#ifdef B200QP_CSL_WITH_DIRECT
    if (c.direct) {   // diagnostic build: every warp reads its tiles with plain global loads (no TMA ring)
      for (int i = warp; i < T; i += CSL_NW) {
        const TileView tv = tile_view<false>(cur.blob + (size_t)__ldg(cur.toff + t0 + i) * 16u, lane);
        body(t0 + i, tv, [] {}, std::false_type{});
      }
      return;
    }
#endif
    const int q0 = ring_q;
    const int qT = (q0 + T) % (2 * S);      // one division per phase
    const int npre = min(KPRE, Tnext);
    if (producer) {
      issue(cur, t0, pre_issued, T, q0);
      issue(nxt, nt0, 0, npre, qT);
    } else {
      int q = q_add(q0, warp);
      for (int i = warp; i < T; i += CSL_NWC) {
        const int stg = q >= S ? q - S : q;
        const unsigned par = q >= S ? 1u : 0u;
        // An mbarrier wait sees one parity bit, so a waiter may be at most ONE phase ahead of the barrier.  The tiles of a
        // stage go to different warps: this warp can come for use k of the stage while the fill of use k-1 (another
        // warp's tile, issued earlier but free to land later) is still in flight -- full[stg] would then still be in
        // phase k-1 and "parity of phase k" would read as completed: the warp would walk into a stale tile.  Waiting first
        // for the RELEASE of use k-1 (same barrier and parity the producer waits on before it refills the stage) pins the
        // full barrier to phase k; that wait is itself unambiguous (release k-2 precedes the issue of use k-1, which
        // precedes every tile this warp has already consumed).  Usually long satisfied: one extra try_wait.
        mbar_wait(&bar_empty[stg], par);
        mbar_wait(&bar_full[stg], par);
        const TileView tv = tile_view<true>(ring + (size_t)stg * c.capb, lane);
        body(t0 + i, tv, [&] { __syncwarp(); if (lane == 0) csl_mbar_arrive(&bar_empty[stg]); }, std::true_type{});
        q = q_add(q, CSL_NWC);
      }
    }
    ring_q = qT;
    pre_issued = npre;
  };
  // before leaving the kernel: the tiles issued ahead must have landed (no bulk copy may target a dead CTA)
  auto ring_drain = [&]() {
    if (producer && lane == 0)
      for (int i = 0; i < pre_issued; ++i) {
        const int q = q_add(ring_q, i);
        mbar_wait(&bar_full[q >= S ? q - S : q], q >= S ? 1u : 0u);
      }
  };
#define CSL_DIE() do { if (cta == 0 && tid == 0) a.out->aborted = 1; return; } while (0)
#define CSL_BARRIER_OR_DIE() do { if (!barrier()) CSL_DIE(); } while (0)

  __shared__ long long t_phase[8];
  __shared__ unsigned long long t_mark;
  if (tid == 0) { for (int i = 0; i < 8; ++i) t_phase[i] = 0; t_mark = global_timer_ns(); }
  auto mark = [&](int id) {
    if (cta == 0 && tid == 0) { const unsigned long long now = global_timer_ns(); t_phase[id] += (long long)(now - t_mark); t_mark = now; }
  };

  // ------------------------------------------------------------------ init: ax = A x, w2 = rho (z - Ax) - y
  ring_phase(c.M, tm0, TM, c.N, tn0, TN, [&](int tile, const TileView& tv, auto&& done, auto smem_tag) {
      constexpr bool SM = decltype(smem_tag)::value;
    double acc, dummy;
    part_dot<1, SM>(tv.i1, tv.v1, tv.len1, tv.max1, a.x, nullptr, acc, dummy);
    done();
    const int row = tile * 32 + lane;
    if (row < a.m) {
      a.ax[row] = acc;
      const double w2n = a.rho_v[row] * (a.z[row] - acc) - a.y[row];
      a.w2[row] = w2n;
      CSL_PUSH(w2, row, w2n);
    }
  });
  CSL_BARRIER_OR_DIE();

  // ------------------------------------------------------------------ R phase
  // epilogue of one n-row: residual, preconditioned start, and on full iterations every norm of the checks
  // hot iterations reduce three values (slots 0..2: <r,u>, |r|inf, clock flag); full iterations the R_K slots of admm_kernel.cuh
  auto r_row = [&](int row, double px, double atw, double aty, auto full_tag, auto& v) {
    constexpr bool FULL = decltype(full_tag)::value;
    const double qj = a.q[row];
    const double rj = -(px + qj) + atw;
    const double uj = a.Minv[row] * rj;
    a.r[row] = rj;
    CSL_PUT_U(row, uj);
    v[R_RZ] += rj * uj;
    v[R_R0] = nanmax(v[R_R0], fabs(rj));
    if constexpr (FULL) {
      a.px[row] = px;
      const double xj = a.x[row], dj = alpha * a.d[row];
      const double dres = px + qj + aty;
      const double di = Dinv_t ? Dinv_t[row] : 1.0;
      v[R_SDUAL] = nanmax(v[R_SDUAL], fabs(dres));
      v[R_SATY] = nanmax(v[R_SATY], fabs(aty));
      v[R_SPX] = nanmax(v[R_SPX], fabs(px));
      v[R_TDUAL] = nanmax(v[R_TDUAL], fabs(di * dres));
      v[R_TATY] = nanmax(v[R_TATY], fabs(di * aty));
      v[R_TPX] = nanmax(v[R_TPX], fabs(di * px));
      v[R_XPX] += xj * px;
      v[R_QX] += qj * xj;
      v[R_NDX] = nanmax(v[R_NDX], fabs((D_t ? D_t[row] : 1.0) * dj));
      v[R_QDX] += qj * dj;
    }
  };
  constexpr int H_TL = 2, H_K = 3;
  constexpr unsigned H_MAXMASK = (1u << R_R0) | (1u << H_TL);
  auto load_R = [&](bool full) {
    if (!full) {
      reduce_load(std::integral_constant<int, H_K>{}, H_MAXMASK);
      rz = s_tot[R_RZ];
      r0max = s_tot[R_R0];
      time_up = s_tot[H_TL] > 0.5;
      return;
    }
    reduce_load(std::integral_constant<int, (int)R_K>{}, R_MAXMASK);
    rz = s_tot[R_RZ];
    r0max = s_tot[R_R0];
    time_up = s_tot[R_TL] > 0.5;
    s_dual = s_tot[R_SDUAL];
    s_ndual = fmax(fmax(s_tot[R_SATY], s_tot[R_SPX]), a.nq_s);
    dual_res = cinv_eff * s_tot[R_TDUAL];
    ndual = cinv_eff * fmax(fmax(s_tot[R_TATY], s_tot[R_TPX]), a.nq_t);
    obj = (0.5 * s_tot[R_XPX] + s_tot[R_QX]) * cinv;
    dobj = (-0.5 * s_tot[R_XPX] - sup) * cinv;
    gap = obj - dobj;
    ndx = s_tot[R_NDX];
    qdx = s_tot[R_QDX];
    prim_res = t_prim;
    nprim = fmax(t_ax, t_z);
  };
  auto phase_R = [&](auto full_tag) {
    constexpr bool FULL = decltype(full_tag)::value;
    if constexpr (DIST) tag_u = x.epoch0 + sync.epoch;
    constexpr int K = FULL ? (int)R_K : H_K;
    constexpr int TL = FULL ? (int)R_TL : H_TL;
    double v[K];
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = 0.0;
    ring_phase(c.N, tn0, TN, c.M, tm0, TM, [&](int tile, const TileView& tv, auto&& done, auto smem_tag) {
      constexpr bool SM = decltype(smem_tag)::value;
      const int row = tile * 32 + lane;
      if constexpr (FULL) {
        // three separate passes: this path runs once per check_termination iterations, registers matter more than latency
        double px, atw, aty, dummy;
        part_dot<1, SM>(tv.i1, tv.v1, tv.len1, tv.max1, a.x, nullptr, px, dummy);
        part_dot<1, SM>(tv.i2, tv.v2, tv.len2, tv.max2, a.w2, nullptr, atw, dummy);
        part_dot<1, SM>(tv.i2, tv.v2, tv.len2, tv.max2, a.y, nullptr, aty, dummy);
        done();
        if (row < a.n) r_row(row, px, atw, aty, full_tag, v);
      } else {
        const double qj = row < a.n ? a.q[row] : 0.0, mi = row < a.n ? a.Minv[row] : 0.0;   // in flight with the gathers
        double px, atw, dummy;
        part_dot<1, SM>(tv.i1, tv.v1, tv.len1, tv.max1, a.x, nullptr, px, dummy);
        part_dot<1, SM>(tv.i2, tv.v2, tv.len2, tv.max2, a.w2, nullptr, atw, dummy);
        done();
        if (row < a.n) {
          const double rj = -(px + qj) + atw;
          const double uj = mi * rj;
          a.r[row] = rj;
          CSL_PUT_U(row, uj);
          v[R_RZ] += rj * uj;
          v[R_R0] = nanmax(v[R_R0], fabs(rj));
        }
      }
    });
    // the clock is sampled every iteration (upstream osqp_solve does; a check_termination = 0 solve must still stop)
    if (tid == 0) v[TL] = ((double)(global_timer_ns() - t_start) > a.time_limit_ns) ? 1.0 : 0.0;
    reduce_store(v, std::integral_constant<int, K>{}, FULL ? R_MAXMASK : H_MAXMASK);
  };
  // after a rho update: px is stored, only A'w2 (and the Jacobi diagonal) change.  Off the alternating ring: plain loads.
  auto phase_R_lite = [&]() {
    if constexpr (DIST) tag_u = x.epoch0 + sync.epoch;
    double v[H_K];
#pragma unroll
    for (int i = 0; i < H_K; ++i) v[i] = 0.0;
    for (int i = warp; i < TN; i += CSL_NW) {
      const int tile = tn0 + i;
      const TileView tv = tile_view<false>(c.N.blob + (size_t)__ldg(c.N.toff + tile) * 16u, lane);
      double atw, dummy, msum = 0.0;
      part_dot<1, false>(tv.i2, tv.v2, tv.len2, tv.max2, a.w2, nullptr, atw, dummy);
      if (a.cg_precond) msum = part_dot_sq<false>(tv.i2, tv.v2, tv.len2, tv.max2, a.rho_v);
      const int row = tile * 32 + lane;
      if (row < a.n) {
        if (a.cg_precond) a.Minv[row] = 1.0 / (a.pdiag[row] + sigma + msum);
        r_row(row, a.px[row], atw, 0.0, std::false_type{}, v);
      }
    }
    if (tid == 0) v[H_TL] = ((double)(global_timer_ns() - t_start) > a.time_limit_ns) ? 1.0 : 0.0;
    reduce_store(v, std::integral_constant<int, H_K>{}, H_MAXMASK);
  };

  // n-rows SpMV + inf-norm for the infeasibility certificates (rare; plain loads): part 1 = P, part 2 = A'
  auto cert_norm = [&](bool use_P, const double* vec, double scale, bool& ok) -> double {
    double v[1] = {0.0};
    for (int i = warp; i < TN; i += CSL_NW) {
      const int tile = tn0 + i;
      const TileView tv = tile_view<false>(c.N.blob + (size_t)__ldg(c.N.toff + tile) * 16u, lane);
      double acc, dummy;
      if (use_P) part_dot<1, false>(tv.i1, tv.v1, tv.len1, tv.max1, vec, nullptr, acc, dummy);
      else part_dot<1, false>(tv.i2, tv.v2, tv.len2, tv.max2, vec, nullptr, acc, dummy);
      const int row = tile * 32 + lane;
      if (row < a.n) v[0] = nanmax(v[0], fabs((Dinv_t ? Dinv_t[row] : 1.0) * scale * acc));
    }
    reduce_store(v, std::integral_constant<int, 1>{}, 1u);
    ok = barrier();
    if (!ok) return 0.0;
    reduce_load(std::integral_constant<int, 1>{}, 1u);
    return s_tot[0];
  };

  // termination test (reference: _osqp.py:998-1077); f = 10 -> "inaccurate"
  auto evaluate = [&](double f, bool& ok) -> int {
    ok = true;
    if (!(prim_res <= QP_INFTY) || !(dual_res <= QP_INFTY)) return 9;
    const double ea = f * a.eps_abs, er = f * a.eps_rel;
    const bool prim_ok = (a.m == 0) || (prim_res < ea + er * nprim);
    const bool dual_ok = dual_res < ea + er * ndual;
    const bool gap_ok = !a.check_dualgap || (fabs(gap) < ea + er * fmax(fabs(obj), fabs(dobj)));
    if (prim_ok && dual_ok && gap_ok) return f > 1.0 ? 2 : 1;
    if (a.polish_mode) return 0;
    if (!prim_ok) {
      const double eps = f * a.eps_pinf;
      if (ndy > eps && lhs < -eps * ndy) {
        const double nat = cert_norm(false, a.w, 1.0, ok);   // w holds the projected delta_y
        if (!ok) return 0;
        if (nat < eps * ndy) return f > 1.0 ? 4 : 3;
      }
    }
    if (!dual_ok) {
      const double eps = f * a.eps_dinf;
      if (ndx > eps && qdx < -c_eff * eps * ndx && vu <= eps * ndx && vl <= eps * ndx) {
        const double npd = cert_norm(true, a.d, alpha, ok);
        if (!ok) return 0;
        if (npd < c_eff * eps * ndx) return f > 1.0 ? 6 : 5;
      }
    }
    return 0;
  };

  // ================================================================== main loop
  for (;;) {
    const bool last = (k >= a.max_iter);
    const bool tlp = tl_pending;      // the clock ran out during the previous iteration: evaluate this iterate, then stop
    const bool chk = k > 0 && ((a.check_term > 0 && (k % a.check_term) == 0) || last || tlp);
    const bool adp = k > 0 && !last && !tlp && a.adaptive_rho && a.adapt_interval > 0 && (k % a.adapt_interval) == 0;
    const bool full = chk || adp;

    {
      if (full) phase_R(std::true_type{}); else phase_R(std::false_type{});
      CSL_BARRIER_OR_DIE();
      load_R(full);
      mark(0);
    }
    if (time_up) tl_pending = true;
    if (full && a.verbose && cta == 0 && tid == 0 && chk && !tlp && (!DIST || x.rank == 0)) {
      const int row = (k / (a.check_term > 0 ? a.check_term : 1)) - 1;
      if (row >= 0 && row < a.log_cap) {
        a.log[row] = LogRow{k, 0, obj, prim_res, dual_res, rho, (double)(global_timer_ns() - t_start) * 1e-6};
        a.out->log_rows = row + 1;
      }
    }
    if (chk) {
      bool ok;
      status = evaluate(1.0, ok);
      if (!ok) CSL_DIE();
      if (status) break;
      if (last) {
        status = evaluate(10.0, ok);
        if (!ok) CSL_DIE();
        if (!status) status = 7;  // MAX_ITER_REACHED
        break;
      }
    }
    if (tlp) { status = 8; break; }   // TIME_LIMIT_REACHED, with the residuals of the returned iterate
    if (adp) {
      // reference: _osqp.py:880-930 (scaled quantities)
      const double pr = s_prim / (s_nprim + 1e-10);
      const double du = s_dual / (s_ndual + 1e-10);
      double rho_new = rho * sqrt(pr / (du + 1e-10));
      rho_new = fmin(fmax(rho_new, RHO_MIN), RHO_MAX);
      if (rho_new > a.adapt_tol * rho || rho_new < rho / a.adapt_tol) {
        rho = rho_new;
        rho_updates++;
        if constexpr (DIST) {
          // every rank re-derives the whole rho vector (the A' rows of its slice gather all of it) ...
          for (int i = um0 + tid; i < um1; i += CSL_NT) {
            const double rv = rho_of_type(a.ctype[i], rho, a.rho_is_vec, a.rho_eq_factor);
            a.rho_v[i] = rv;
            a.rho_inv[i] = 1.0 / rv;
          }
          // ... and refreshes w2 on its own rows for everybody
          for (int i = m0 + tid; i < m1; i += CSL_NT) {
            const double rv = rho_of_type(a.ctype[i], rho, a.rho_is_vec, a.rho_eq_factor);
            const double w2n = rv * (a.z[i] - a.ax[i]) - a.y[i];
            a.w2[i] = w2n;
            CSL_PUSH(w2, i, w2n);
          }
        } else {
          for (int i = m0 + tid; i < m1; i += CSL_NT) {
            const double rv = rho_of_type(a.ctype[i], rho, a.rho_is_vec, a.rho_eq_factor);
            a.rho_v[i] = rv;
            a.rho_inv[i] = 1.0 / rv;
            a.w2[i] = rv * (a.z[i] - a.ax[i]) - a.y[i];
          }
        }
        CSL_BARRIER_OR_DIE();
        phase_R_lite();
        CSL_BARRIER_OR_DIE();
        load_R(false);
      }
    }

    // ---------------------------------------------------------------- PCG on K d = r (single-reduction recurrence)
    // a.p holds u = M^-1 r and rz = <r, u> from the R phase; a.kp holds w = K u; ps / ss the direction and K * direction
    bool have_d = false;
    if (r0max > 1e-150) {
      const double tol = a.cg_tol_frac * r0max;
      double gam = rz, gam_prev = rz, alpha_cg = 0.0;
      for (int j = 0; j < a.cg_max_iter; ++j) {
        // B: t = rho .* (A u)
        const unsigned long long tb0 = (c.cta_ns && tid == 0) ? global_timer_ns() : 0ULL;
        ring_phase(c.M, tm0, TM, c.N, tn0, TN, [&](int tile, const TileView& tv, auto&& done, auto smem_tag) {
      constexpr bool SM = decltype(smem_tag)::value;
          const int row = tile * 32 + lane;
          const double rv = row < a.m ? __ldcg(a.rho_v + row) : 0.0;      // in flight with the gathers
          double acc, dummy;
          bool tg_u = false;
          if constexpr (DIST && SM) tg_u = tagged_u;
          if (tg_u) part_dot_tag(tv.i1, tv.v1, tv.len1, tv.max1, tu_mine, tag_u, acc);
          else part_dot<1, SM>(tv.i1, tv.v1, tv.len1, tv.max1, a.p, nullptr, acc, dummy);
          done();
          if (row < a.m) {
            const double tv_ = rv * acc;
            if constexpr (DIST) {
              if (x.tag_t) {      // {t, phase tag} to every rank's tagged copy (own included): C gathers it without a barrier
                const unsigned long long tg = x.epoch0 + sync.epoch;
#pragma unroll
                for (int h = 0; h < CSL_MAXW; ++h) if (h < x.world) csl_st_slot(x.tt[h] + (size_t)row * 16, tv_, tg);
              } else { a.w[row] = tv_; CSL_PUSH(w, row, tv_); }
            } else {
              a.w[row] = tv_;
            }
          }
        });
        if (c.cta_ns) { __syncthreads(); if (tid == 0) c.cta_ns[2 * cta] += (long long)(global_timer_ns() - tb0); }
        if (!(DIST && x.tag_t)) CSL_BARRIER_OR_DIE();
        mark(1);
        const unsigned long long tc0 = (c.cta_ns && tid == 0) ? global_timer_ns() : 0ULL;
        // C: w = P u + sigma u + A' t ; <w, u>
        double v1[1] = {0.0};
        ring_phase(c.N, tn0, TN, c.M, tm0, TM, [&](int tile, const TileView& tv, auto&& done, auto smem_tag) {
      constexpr bool SM = decltype(smem_tag)::value;
          const int row = tile * 32 + lane;
          const double uj = row < a.n ? __ldcg(a.p + row) : 0.0;          // in flight with the gathers
          double accP, accA, dummy;
          bool tg_u = false;
          if constexpr (DIST && SM) tg_u = tagged_u;
          if (tg_u) part_dot_tag(tv.i1, tv.v1, tv.len1, tv.max1, tu_mine, tag_u, accP);
          else part_dot<1, SM>(tv.i1, tv.v1, tv.len1, tv.max1, a.p, nullptr, accP, dummy);
          bool tagged = false;
          if constexpr (DIST && SM) tagged = x.tag_t != 0;
          if (tagged) {
            if constexpr (DIST) {
              const unsigned char* mine = x.tt[0];
#pragma unroll
              for (int h = 1; h < CSL_MAXW; ++h) if (x.rank == h) mine = x.tt[h];
              part_dot_tag(tv.i2, tv.v2, tv.len2, tv.max2, mine, x.epoch0 + sync.epoch, accA);
            }
          } else {
            part_dot<1, SM>(tv.i2, tv.v2, tv.len2, tv.max2, a.w, nullptr, accA, dummy);
          }
          done();
          if (row < a.n) {
            const double wj = (accP + accA) + sigma * uj;
            a.kp[row] = wj;
            v1[0] += uj * wj;
          }
        });
        if (c.cta_ns && tid == 0) c.cta_ns[2 * cta + 1] += (long long)(global_timer_ns() - tc0);
        if (!barrier_payload(v1, std::integral_constant<int, 1>{}, 0u, 0)) CSL_DIE();
        mark(2);
        const double delta = s_tot[0];
        const double bj = (j == 0) ? 0.0 : gam / gam_prev;
        const double pkp = (j == 0) ? delta : delta - bj * gam / alpha_cg;      // = <p, K p>
        if (!(pkp > 0.0)) { status = 9; break; }   // K not positive definite (or NaN): non-convex
        alpha_cg = gam / pkp;
        cg_total++;
        // V: p = u + b p ; s = w + b s ; d += a p ; r -= a s ; u = Minv r ; <r, u> ; |r|inf   (two elements per thread)
        double v2[2] = {0.0, 0.0};
        const unsigned long long tag_v = DIST ? x.epoch0 + sync.epoch : 0ULL;
        {
          const int npair = (n1 - n0) >> 1;
          const double2* u2 = reinterpret_cast<const double2*>(a.p + n0);
          const double2* w2v = reinterpret_cast<const double2*>(a.kp + n0);
          const double2* mi2 = reinterpret_cast<const double2*>(a.Minv + n0);
          double2* p2 = reinterpret_cast<double2*>(c.ps + n0);
          double2* s2 = reinterpret_cast<double2*>(c.ss + n0);
          double2* d2 = reinterpret_cast<double2*>(a.d + n0);
          double2* r2 = reinterpret_cast<double2*>(a.r + n0);
          double2* uo2 = reinterpret_cast<double2*>(a.p + n0);
          auto upd = [&](double ui, double wi, double po, double so, double dold, double ro, double mi, double& pn, double& sn,
                         double& dn, double& rn, double& un) {
            pn = (j == 0) ? ui : fma(bj, po, ui);
            sn = (j == 0) ? wi : fma(bj, so, wi);
            dn = (j == 0) ? alpha_cg * pn : fma(alpha_cg, pn, dold);
            rn = fma(-alpha_cg, sn, ro);
            un = mi * rn;
            v2[0] += rn * un;
            v2[1] = nanmax(v2[1], fabs(rn));
          };
          for (int i = tid; i < npair; i += CSL_NT) {
            const double2 ui = __ldcg(u2 + i), wi = __ldcg(w2v + i), mi = __ldg(mi2 + i), ro = __ldcg(r2 + i);
            double2 po = make_double2(0.0, 0.0), so = po, dold = po;
            if (j > 0) { po = __ldcg(p2 + i); so = __ldcg(s2 + i); dold = __ldcg(d2 + i); }
            double2 pn, sn, dn, rn, un;
            upd(ui.x, wi.x, po.x, so.x, dold.x, ro.x, mi.x, pn.x, sn.x, dn.x, rn.x, un.x);
            upd(ui.y, wi.y, po.y, so.y, dold.y, ro.y, mi.y, pn.y, sn.y, dn.y, rn.y, un.y);
            p2[i] = pn; s2[i] = sn; d2[i] = dn; r2[i] = rn; uo2[i] = un;
            if constexpr (DIST) {
              if (tagged_u) {
#pragma unroll
                for (int h_ = 0; h_ < CSL_MAXW; ++h_)
                  if (h_ < x.world) {
                    csl_st_slot(x.tu[h_] + (size_t)(n0 + 2 * i) * 16, un.x, tag_v);
                    csl_st_slot(x.tu[h_] + (size_t)(n0 + 2 * i + 1) * 16, un.y, tag_v);
                  }
              } else {
                CSL_PUSH2(p, (n0 >> 1) + i, un);
              }
            }
          }
          if (((n1 - n0) & 1) && tid == 0) {
            const int i = n1 - 1;
            double pn, sn, dn, rn, un;
            upd(__ldcg(a.p + i), __ldcg(a.kp + i), j > 0 ? c.ps[i] : 0.0, j > 0 ? c.ss[i] : 0.0, j > 0 ? a.d[i] : 0.0, a.r[i],
                a.Minv[i], pn, sn, dn, rn, un);
            c.ps[i] = pn; c.ss[i] = sn; a.d[i] = dn; a.r[i] = rn; a.p[i] = un;
            if constexpr (DIST) {
              if (tagged_u) {
#pragma unroll
                for (int h_ = 0; h_ < CSL_MAXW; ++h_) if (h_ < x.world) csl_st_slot(x.tu[h_] + (size_t)i * 16, un, tag_v);
              } else {
                CSL_PUSH(p, i, un);
              }
            }
          }
        }
        have_d = true;
        if constexpr (DIST) tag_u = tag_v;          // the gathers of the next A u / P u expect this phase's u
        if (tagged_u) {
          // nothing of this phase has to have landed anywhere when the next one starts: fence-free payload barrier
          if (!barrier_payload(v2, std::integral_constant<int, 2>{}, 2u, 1)) CSL_DIE();
        } else {
          reduce_store(v2, std::integral_constant<int, 2>{}, 2u);
          CSL_BARRIER_OR_DIE();
          reduce_load(std::integral_constant<int, 2>{}, 2u);
        }
        mark(3);
        gam_prev = gam;
        gam = s_tot[0];
        const double rmax = s_tot[1];
        if (!(rmax > tol) || j + 1 == a.cg_max_iter) break;
      }
      if (status) break;
    }
    if constexpr (DIST) {
      // the Z phase gathers from the whole step: all-gather d (every rank produced its slice), or zero every copy
      if (have_d) {
        const int npair = (n1 - n0) >> 1;
        const double2* d2 = reinterpret_cast<const double2*>(a.d + n0);
        for (int i = tid; i < npair; i += CSL_NT) { const double2 v = d2[i]; CSL_PUSH2(d, (n0 >> 1) + i, v); }
        if (((n1 - n0) & 1) && tid == 0) { const double v = a.d[n1 - 1]; CSL_PUSH(d, n1 - 1, v); }
      } else {
        for (int i = un0 + tid; i < un1; i += CSL_NT) a.d[i] = 0.0;
      }
      CSL_BARRIER_OR_DIE();
      mark(6);
    } else if (!have_d) {
      for (int i = n0 + tid; i < n1; i += CSL_NT) a.d[i] = 0.0;
      CSL_BARRIER_OR_DIE();
    }

    // ---------------------------------------------------------------- Z phase
    const int kn = k + 1;
    const bool last_n = (kn >= a.max_iter);
    const bool full_n = ((a.check_term > 0 && (kn % a.check_term) == 0) || last_n || tl_pending) ||
                        (!last_n && a.adaptive_rho && a.adapt_interval > 0 && (kn % a.adapt_interval) == 0);
    auto phase_Z = [&](auto full_tag) -> bool {
      constexpr bool FULL = decltype(full_tag)::value;
      double v[FULL ? (int)Z_K : 1];
#pragma unroll
      for (int i = 0; i < (FULL ? (int)Z_K : 1); ++i) v[i] = 0.0;
      if constexpr (FULL) { v[Z_VU] = NEG_INF; v[Z_VL] = NEG_INF; }
      auto z_row = [&](int row, double ad) {
        const double ax_old = a.ax[row], z_old = a.z[row], y_old = a.y[row];
        const double rv = a.rho_v[row], ri = a.rho_inv[row], li = a.l[row], ui = a.u[row];
        const double zt = ax_old + ad;                                   // z~ = A x~
        const double zr = alpha * zt + (1.0 - alpha) * z_old;            // _osqp.py:686-690
        const double zn = fmin(fmax(zr + ri * y_old, li), ui);           // _osqp.py:670-674
        const double dy = rv * (zr - zn);                                // _osqp.py:698-703
        const double yn = y_old + dy;
        const double axn = fma(alpha, ad, ax_old);
        a.z[row] = zn; a.y[row] = yn; a.ax[row] = axn;
        const double w2n = rv * (zn - axn) - yn;
        a.w2[row] = w2n;
        CSL_PUSH(w2, row, w2n);
        if constexpr (FULL) {
          CSL_PUSH(y, row, yn);
          const double ei = Einv_t ? Einv_t[row] : 1.0;
          const double pr = axn - zn;
          v[Z_SPRIM] = nanmax(v[Z_SPRIM], fabs(pr));
          v[Z_SAX] = nanmax(v[Z_SAX], fabs(axn));
          v[Z_SZ] = nanmax(v[Z_SZ], fabs(zn));
          v[Z_TPRIM] = nanmax(v[Z_TPRIM], fabs(ei * pr));
          v[Z_TAX] = nanmax(v[Z_TAX], fabs(ei * axn));
          v[Z_TZ] = nanmax(v[Z_TZ], fabs(ei * zn));
          // primal infeasibility (reference _osqp.py:796-820), delta_y projected on the polar of the recession cone of [l,u]
          const bool lfin = li > -INF_THRESH, ufin = ui < INF_THRESH;
          double dyp = dy;
          if (!ufin && !lfin) dyp = 0.0;
          else if (!ufin) dyp = fmin(dy, 0.0);
          else if (!lfin) dyp = fmax(dy, 0.0);
          a.w[row] = dyp;
          CSL_PUSH(w, row, dyp);
          v[Z_NDY] = nanmax(v[Z_NDY], fabs((E_t ? E_t[row] : 1.0) * dyp));
          v[Z_LHS] += (ufin ? ui * fmax(dyp, 0.0) : 0.0) + (lfin ? li * fmin(dyp, 0.0) : 0.0);
          // dual infeasibility rows (reference _osqp.py:861-872)
          const double adx = ei * alpha * ad;
          if (ufin) v[Z_VU] = nanmax(v[Z_VU], adx);
          if (lfin) v[Z_VL] = nanmax(v[Z_VL], -adx);
          v[Z_SUP] += (ufin ? ui * fmax(yn, 0.0) : 0.0) + (lfin ? li * fmin(yn, 0.0) : 0.0);
        }
      };
      ring_phase(c.M, tm0, TM, c.N, tn0, TN, [&](int tile, const TileView& tv, auto&& done, auto smem_tag) {
      constexpr bool SM = decltype(smem_tag)::value;
        double ad, dummy;
        part_dot<1, SM>(tv.i1, tv.v1, tv.len1, tv.max1, a.d, nullptr, ad, dummy);
        done();
        const int row = tile * 32 + lane;
        if (row < a.m) z_row(row, ad);
      });
      for (int i = un0 + tid; i < un1; i += CSL_NT) a.x[i] = fma(alpha, __ldcg(a.d + i), a.x[i]);  // _osqp.py:664-668
      if constexpr (FULL) {
        reduce_store(v, std::integral_constant<int, (int)Z_K>{}, Z_MAXMASK);
        if (!barrier()) return false;
        reduce_load(std::integral_constant<int, (int)Z_K>{}, Z_MAXMASK);
        s_prim = s_tot[Z_SPRIM];
        s_nprim = fmax(s_tot[Z_SAX], s_tot[Z_SZ]);
        t_prim = s_tot[Z_TPRIM]; t_ax = s_tot[Z_TAX]; t_z = s_tot[Z_TZ];
        ndy = s_tot[Z_NDY]; lhs = s_tot[Z_LHS]; vu = s_tot[Z_VU]; vl = s_tot[Z_VL]; sup = s_tot[Z_SUP];
        return true;
      } else {
        return barrier();
      }
    };
    if (!(full_n ? phase_Z(std::true_type{}) : phase_Z(std::false_type{}))) CSL_DIE();
    mark(5);
    k = kn;
  }

  ring_drain();
  if (cta == 0 && tid == 0) {
    KOut* o = a.out;
    for (int i = 0; i < 8; ++i) o->phase_ns[i] = t_phase[i];
    o->status = status; o->iter = k; o->rho_updates = rho_updates;
    o->cg_iters = cg_total; o->barriers = (long long)sync.epoch;
    if constexpr (DIST) o->xepoch = (long long)(x.epoch0 + sync.epoch);
    o->obj = obj; o->dual_obj = dobj; o->gap = gap; o->prim_res = prim_res; o->dual_res = dual_res;
    o->rho = rho;
    const double pr = s_prim / (s_nprim + 1e-10), du = s_dual / (s_ndual + 1e-10);
    o->rho_estimate = fmin(fmax(rho * sqrt(pr / (du + 1e-10)), RHO_MIN), RHO_MAX);
    o->rel_kkt = fmax(prim_res / (nprim + 1e-10), dual_res / (ndual + 1e-10));
  }
#undef CSL_DIE
#undef CSL_BARRIER_OR_DIE
#undef CSL_PUSH
#undef CSL_PUSH2
#undef CSL_PUT_U
}

}  // namespace b200qp
