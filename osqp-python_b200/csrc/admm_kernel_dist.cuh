// admm_kernel_dist.cuh -- the persistent ADMM kernel for ONE QP partitioned over several GPUs.
//
// Partition.  Rank g owns
//   * a contiguous slab of constraint rows: A_g, and z, y, l, u, rho, A x on those rows          (m-side work)
//   * a contiguous slice of the variables: the rows of P and of A' in that slice, and r, Kp, M^-1 (n-side work)
// so every SpMV and every vector pass shrinks with the number of GPUs.  What each rank needs in full is what the
// other side gathers from: p and d (n-vectors, gathered by A_g rows) and w, w2, y (m-vectors, gathered by A'
// rows).  Those five vectors live in a peer-mapped region; a rank writes its own slab / slice locally and PUSHES it
// into every peer's copy with st.relaxed.sys over NVLink / NVSwitch right where it is produced (all-gather by
// stores, fused into the producing phase -- no NCCL call, no extra launch, no host in the loop); x is kept
// replicated by applying the same update from the same d on every rank.  Ranks meet at cross-GPU barriers built on
// flags in peer-mapped memory; the barrier carries the scalar partials (dots, norms) as payload, folded in rank
// order, so all ranks see identical bits and take identical control decisions.
//
// The PCG is the single-reduction variant of Chronopoulos & Gear (same iterates as the textbook recurrence in exact
// arithmetic: <p,Kp> = <w,u> - beta <r,u> / alpha_prev with u = M^-1 r, w = K u), because every global reduction is a
// cross-GPU barrier here.  Per PCG iteration: B (t = rho A_g u -> push t) | C (slice rows: w = P u + sigma u + A' t,
// <w,u>) | V (slice: p, s, d, r, u = M^-1 r -> push u, <r,u>, |r|inf): three cross-GPU barriers, 2 x 8 (G-1)/G MB
// pushed per rank, no replicated work.
// (The first version reduce-scattered partial A_g' w products of full length n: every rank walked all n rows of a
// matrix with 10/G entries per row and repeated the C3 / C4 passes -- it stopped scaling at 2 GPUs, profiles/r1d.)
#pragma once
#include "admm_kernel.cuh"

namespace b200qp {

constexpr int MAXW = 8;

struct XArgs {
  int rank, world;
  int slice[MAXW + 1];              // variable-space slice boundaries
  int mlo, mtot;                    // first row of my slab, total number of constraint rows
  // peer-mapped tables: entry h points into rank h's shared region
  double* wf[MAXW];                 // full m-vector w  (rho .* A p; the projected delta_y on check iterations)
  double* w2f[MAXW];                // full m-vector w2 (rho (z - A x) - y)
  double* yf[MAXW];                 // full m-vector y  (refreshed on check iterations only)
  double* p[MAXW];                  // full n-vector p
  double* d[MAXW];                  // full n-vector d
  unsigned long long* flags[MAXW];  // rank h's arrival counters, one per source rank
  double* payload[MAXW];            // rank h's payload area [2][MAXW][KMAX]
  int* abort_flag[MAXW];
  // local
  double* xtot;                     // [2][KMAX] totals published by CTA 0 to the other CTAs
  unsigned long long* release;      // local release counter of the cross-GPU barrier
  const int *part_slice, *part_uni, *part_mall;   // CTA partitions: my slice (cost of P + A' rows), uniform [0,n), uniform [0,mtot)
  double *ps, *ss;                  // slice-local search direction p and s = K p of the single-reduction PCG
  const int* ctype_full;            // full-length typing / rho vectors: every rank re-derives all of rho at a rho update
  double *rho_full, *rho_inv_full;
  unsigned long long xepoch0;       // cross-GPU barriers completed by earlier launches (flags are never reset)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_sys_f64(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// Local grid barrier + cross-GPU barrier, optionally all-reducing K scalars.
//   every CTA : block partials already in `bank` (if K) ; arrive
//   CTA 0     : wait for local arrivals, fold the local partials, publish them to every rank with an
//               arrival flag, wait for every rank, fold all ranks' payloads in rank order, release
//   every CTA : totals in s_tot[0..K)
struct DistSync {
  GridSync* g;
  const XArgs* x;
  unsigned long long xepoch;     // cross-GPU barriers passed

  // pushed = this CTA issued peer stores since the last barrier (they must be visible system-wide before the
  // arrival flag); phases that only produce local data and a payload get away with a device-scope fence
  template <int K>
  __device__ __forceinline__ bool barrier_x(const double* bank, unsigned maxmask, double* s_tot, bool pushed = true) {
    const int G = g->G, world = x->world, tid = threadIdx.x;
    if (pushed) __threadfence_system(); else __threadfence();
    __syncthreads();
    g->epoch++;
    xepoch++;
    const int par = (int)(xepoch & 1ULL);
    bool ok = true;
    if (blockIdx.x == 0) {
      // 1. wait for every local CTA
      if (tid == 0 && G == 1) *g->s_ok = 1;
      if (tid == 0 && G > 1) {
        __threadfence();
        atomicAdd(&g->ctrl->bar, 1ULL);
        const unsigned long long target = g->epoch * (unsigned long long)G;
        unsigned spins = 0; unsigned long long t0 = 0;
        while (ld_acquire_u64(&g->ctrl->bar) < target) {
          if ((++spins & 0xfffu) == 0) {
            if (*(volatile int*)&g->ctrl->abort) { ok = false; break; }
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now; else if (now - t0 > 10000000000ULL) { atomicExch(&g->ctrl->abort, 1); ok = false; break; }
          }
        }
        __threadfence();
        *g->s_ok = ok ? 1 : 0;
      }
      __syncthreads();
      ok = *g->s_ok != 0;
      // 2. local totals
      if (K > 0 && ok) grid_reduce_load<(K > 0 ? K : 1)>(bank, G, maxmask, s_tot);
      // 3. publish payload + arrival flag to every rank (thread h talks to rank h)
      if (ok && tid < world) {
        double* dst = x->payload[tid] + ((size_t)par * MAXW + x->rank) * KMAX;
        for (int k = 0; k < K; ++k) st_sys_f64(dst + k, s_tot[k]);
        __threadfence_system();
        st_release_sys_u64(x->flags[tid] + x->rank, xepoch);
      }
      // 4. wait for every rank's arrival (their flags live in my memory)
      if (ok && tid < world) {
        unsigned spins = 0; unsigned long long t0 = 0;
        const unsigned long long* f = x->flags[x->rank] + tid;
        while (ld_acquire_sys_u64(f) < xepoch) {
          if ((++spins & 0x3ffu) == 0) {
            if (*(volatile int*)&g->ctrl->abort) { ok = false; break; }
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now; else if (now - t0 > 10000000000ULL) { atomicExch(&g->ctrl->abort, 1); ok = false; break; }
          }
        }
        if (!ok) *g->s_ok = 0;
      }
      __syncthreads();
      ok = *g->s_ok != 0;
      // 5. fold the payloads of all ranks in rank order, publish to the local CTAs, release
      if (ok && tid < K) {
        const bool mx = (maxmask >> tid) & 1u;
        const double* src = x->payload[x->rank] + (size_t)par * MAXW * KMAX;
        double acc = __ldcv(src + tid);
        for (int h = 1; h < world; ++h) acc = combine(acc, __ldcv(src + (size_t)h * KMAX + tid), mx);
        x->xtot[par * KMAX + tid] = acc;
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        if (!ok) atomicExch(&g->ctrl->abort, 1);
        __threadfence();
        atomicExch(x->release, g->epoch);     // monotone: release value = local epoch
      }
    } else {
      if (tid == 0) {
        __threadfence();
        atomicAdd(&g->ctrl->bar, 1ULL);
        unsigned spins = 0; unsigned long long t0 = 0;
        int good = 1;
        while (ld_acquire_u64(x->release) < g->epoch) {
          if ((++spins & 0xfffu) == 0) {
            if (*(volatile int*)&g->ctrl->abort) { good = 0; break; }
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now; else if (now - t0 > 12000000000ULL) { atomicExch(&g->ctrl->abort, 1); good = 0; break; }
          }
        }
        __threadfence();
        if (*(volatile int*)&g->ctrl->abort) good = 0;
        *g->s_ok = good;
      }
      __syncthreads();
      ok = *g->s_ok != 0;
    }
    if (ok && K > 0) {
      if (tid < K) s_tot[tid] = __ldcg(x->xtot + par * KMAX + tid);
      __syncthreads();
    }
    if (!ok) {   // tell the other ranks so that nobody spins until the watchdog
      if (blockIdx.x == 0 && tid < world) *(volatile int*)x->abort_flag[tid] = 1;
    }
    return ok;
  }
};

#define B200QP_XBAR_OR_DIE(K, bank, mask) do { if (!xs.template barrier_x<K>(bank, mask, s_tot, true)) { if (blockIdx.x == 0 && threadIdx.x == 0) a.out->aborted = 1; return; } } while (0)

template <int LGN, int LGM>
__global__ void __launch_bounds__(NT, 2) admm_kernel_dist(const __grid_constant__ KArgs a, const __grid_constant__ XArgs x) {
  __shared__ double s_red[KMAX * NW];
  __shared__ double s_tot[KMAX];
  __shared__ int s_ok;

  const int G = a.G, cta = blockIdx.x, tid = threadIdx.x;
  const int rank = x.rank, world = x.world, mlo = x.mlo;
  const int s0 = x.part_slice[cta], s1 = x.part_slice[cta + 1];     // my share of this rank's variable slice
  const int u0 = x.part_uni[cta], u1 = x.part_uni[cta + 1];         // uniform share of [0, n) (replicated x)
  const int g0 = x.part_mall[cta], g1 = x.part_mall[cta + 1];       // uniform share of [0, mtot) (rho re-typing)
  const int m0 = a.part_m[cta], m1 = a.part_m[cta + 1];             // my share of this rank's row slab (local rows)
  GridSync sync{a.ctrl, G, 0ULL, &s_ok};
  DistSync xs{&sync, &x, x.xepoch0};
  const double NEG_INF = __longlong_as_double(0xfff0000000000000LL);

  constexpr int NGN = NT / LGN, NGM = NT / LGM;
  const int lane_n = tid & (LGN - 1), grp_n = tid / LGN;
  const int lane_m = tid & (LGM - 1), grp_m = tid / LGM;

  const double alpha = a.alpha, sigma = a.sigma;
  const double cinv = 1.0 / a.c;
  const double c_eff = a.unscale ? a.c : 1.0, cinv_eff = a.unscale ? cinv : 1.0;
  const double* Dinv_t = a.unscale ? a.Dinv : nullptr;
  const double* D_t = a.unscale ? a.D : nullptr;
  const double* Einv_t = a.unscale ? a.Einv : nullptr;
  const double* E_t = a.unscale ? a.E : nullptr;

  // my copies of the gathered vectors (full length); a.w / a.w2 / a.y are the slab views, a.p / a.d the full ones
  double* const wfull = x.wf[rank];
  double* const w2full = x.w2f[rank];
  double* const yfull = x.yf[rank];

  double rho = a.rho;
  int rho_updates = 0, status = 0, k = 0;
  long long cg_total = 0;
  const unsigned long long t_start = global_timer_ns();

  double obj = 0, dobj = 0, gap = 0, prim_res = 0, dual_res = 0, nprim = 0, ndual = 0;
  double s_prim = 0, s_nprim = 0, s_dual = 0, s_ndual = 0;
  double ndy = 0, lhs = 0, vu = NEG_INF, vl = NEG_INF, sup = 0;
  double t_prim = 0, t_ax = 0, t_z = 0;
  double ndx = 0, qdx = 0;
  double rz = 0, r0max = 0;
  bool time_up = false, tl_pending = false;

  auto bank = [&]() -> double* { return a.red + (size_t)(sync.epoch & 1ULL) * KMAX * G; };

  // all-gather by stores: push this CTA's rows [lo, hi) (global indices) of my copy into every peer's copy
  auto push_rows = [&](double* const* table, const double* mine, int lo, int hi) {
    for (int i = lo + tid; i < hi; i += NT) {
      const double v = __ldcg(mine + i);
#pragma unroll
      for (int h = 0; h < MAXW; ++h)
        if (h < world && h != rank) st_sys_f64(table[h] + i, v);
    }
  };

  // one value of a gathered vector, pushed by the thread that produced it: the NVLink stores then overlap the rest
  // of the phase instead of queueing up in front of the barrier's system-scope fence
  auto push_one = [&](double* const* table, int i, double v) {
#pragma unroll
    for (int h = 0; h < MAXW; ++h)
      if (h < world && h != rank) st_sys_f64(table[h] + i, v);
  };

  // ------------------------------------------------------------------ init: ax = A_g x, w2 (slab) -> everybody
  for (int base = m0; base < m1; base += NGM) {
    const int row = base + grp_m;
    const bool valid = row < m1;
    const double acc = group_sum<LGM>(row_dot<LGM>(a.A, a.x, row, valid, lane_m));
    if (valid && lane_m == 0) {
      a.ax[row] = acc;
      a.w2[row] = __ldcg(a.rho_v + row) * (a.z[row] - acc) - a.y[row];
    }
  }
  __syncthreads();
  push_rows(x.w2f, w2full, mlo + m0, mlo + m1);
  B200QP_XBAR_OR_DIE(0, nullptr, 0u);

  __shared__ long long t_phase[8];
  __shared__ unsigned long long t_mark;
  if (tid == 0) { for (int i = 0; i < 8; ++i) t_phase[i] = 0; t_mark = global_timer_ns(); }
  auto mark = [&](int id) {
    if (cta == 0 && tid == 0) { const unsigned long long now = global_timer_ns(); t_phase[id] += (long long)(now - t_mark); t_mark = now; }
  };

  // ---- R phase on my slice: r = -(P x + q) + A' w2, p = M^-1 r (pushed); full = also the dual-side quantities;
  // lite = after a rho update (px kept, Jacobi diagonal from the full rho vector)
  auto phase_R = [&](bool full, bool lite) -> bool {
    double v[R_K];
#pragma unroll
    for (int i = 0; i < R_K; ++i) v[i] = 0.0;
    for (int base = s0; base < s1; base += NGN) {
      const int row = base + grp_n;
      const bool valid = row < s1;
      double px = 0, atw = 0, aty = 0, msum = 0;
      if (lite) {
        atw = row_dot<LGN>(a.At, w2full, row, valid, lane_n);
        if (a.cg_precond) msum = row_dot_sq<LGN>(a.At, x.rho_full, row, valid, lane_n);
        msum = group_sum<LGN>(msum);
      } else {
        if (full) {
          px = row_dot<LGN>(a.P, a.x, row, valid, lane_n);
          row_dot2<LGN>(a.At, w2full, yfull, row, valid, lane_n, atw, aty);
        } else {
          dot_pair<LGN, 1, 4>(a.P, a.x, row, valid, a.At, w2full, row, valid, lane_n, px, atw);
        }
        px = group_sum<LGN>(px);
        if (full) aty = group_sum<LGN>(aty);
      }
      atw = group_sum<LGN>(atw);
      if (valid && lane_n == 0) {
        const double qj = a.q[row];
        if (lite) {
          px = a.px[row];
          if (a.cg_precond) a.Minv[row] = 1.0 / (a.pdiag[row] + sigma + msum);
        }
        const double rj = -(px + qj) + atw;
        const double pj = a.Minv[row] * rj;
        a.r[row] = rj;
        a.p[row] = pj;
        push_one(x.p, row, pj);
        v[R_RZ] += rj * pj;
        v[R_R0] = nanmax(v[R_R0], fabs(rj));
        if (full) {
          a.px[row] = px;
          const double xj = __ldcg(a.x + row), dj = alpha * __ldcg(a.d + row);
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
      }
    }
    if (tid == 0)   // every iteration (see admm_kernel.cuh)
      v[R_TL] = ((double)(global_timer_ns() - t_start) > a.time_limit_ns) ? 1.0 : 0.0;
    double* bk = bank();
    block_reduce_store<R_K>(v, R_MAXMASK, bk, G, s_red);
    if (!xs.template barrier_x<R_K>(bk, R_MAXMASK, s_tot)) return false;
    rz = s_tot[R_RZ];
    r0max = s_tot[R_R0];
    time_up = s_tot[R_TL] > 0.5;
    if (full) {
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
    }
    return true;
  };

  // |Dinv .* scale .* (M vec)|inf over my slice rows, max over ranks (infeasibility certificates)
  auto cert_norm = [&](const CsrMat& M, const double* vec, double scale, bool& ok) -> double {
    double v[1] = {0.0};
    for (int base = s0; base < s1; base += NGN) {
      const int row = base + grp_n;
      const bool valid = row < s1;
      const double acc = group_sum<LGN>(row_dot<LGN>(M, vec, row, valid, lane_n));
      if (valid && lane_n == 0) v[0] = nanmax(v[0], fabs((Dinv_t ? Dinv_t[row] : 1.0) * scale * acc));
    }
    double* bk = bank();
    block_reduce_store<1>(v, 1u, bk, G, s_red);
    ok = xs.template barrier_x<1>(bk, 1u, s_tot, false);
    return ok ? s_tot[0] : 0.0;
  };

  auto evaluate = [&](double f, bool& ok) -> int {
    ok = true;
    if (!(prim_res <= QP_INFTY) || !(dual_res <= QP_INFTY)) return 9;
    const double ea = f * a.eps_abs, er = f * a.eps_rel;
    const bool prim_ok = (x.mtot == 0) ? true : (prim_res < ea + er * nprim);
    const bool dual_ok = dual_res < ea + er * ndual;
    const bool gap_ok = !a.check_dualgap || (fabs(gap) < ea + er * fmax(fabs(obj), fabs(dobj)));
    if (prim_ok && dual_ok && gap_ok) return f > 1.0 ? 2 : 1;
    if (a.polish_mode) return 0;
    if (!prim_ok) {
      const double eps = f * a.eps_pinf;
      if (ndy > eps && lhs < -eps * ndy) {
        const double nat = cert_norm(a.At, wfull, 1.0, ok);     // w holds the projected delta_y (pushed by the Z phase)
        if (!ok) return 0;
        if (nat < eps * ndy) return f > 1.0 ? 4 : 3;
      }
    }
    if (!dual_ok) {
      const double eps = f * a.eps_dinf;
      if (ndx > eps && qdx < -c_eff * eps * ndx && vu <= eps * ndx && vl <= eps * ndx) {
        const double npd = cert_norm(a.P, a.d, alpha, ok);
        if (!ok) return 0;
        if (npd < c_eff * eps * ndx) return f > 1.0 ? 6 : 5;
      }
    }
    return 0;
  };

#define B200QP_XBAR_LOCAL_OR_DIE(K, bank, mask) do { if (!xs.template barrier_x<K>(bank, mask, s_tot, false)) { if (blockIdx.x == 0 && threadIdx.x == 0) a.out->aborted = 1; return; } } while (0)
#define B200QP_DIE() do { if (cta == 0 && tid == 0) a.out->aborted = 1; return; } while (0)

  // ================================================================== main loop
  for (;;) {
    const bool last = (k >= a.max_iter);
    const bool tlp = tl_pending;
    const bool chk = k > 0 && ((a.check_term > 0 && (k % a.check_term) == 0) || last || tlp);
    const bool adp = k > 0 && !last && !tlp && a.adaptive_rho && a.adapt_interval > 0 && (k % a.adapt_interval) == 0;
    const bool full = chk || adp;

    if (!phase_R(full, false)) B200QP_DIE();
    mark(0);
    if (time_up) tl_pending = true;
    if (full && a.verbose && cta == 0 && tid == 0 && chk && !tlp && rank == 0) {
      const int row = (k / (a.check_term > 0 ? a.check_term : 1)) - 1;
      if (row >= 0 && row < a.log_cap) {
        a.log[row] = LogRow{k, 0, obj, prim_res, dual_res, rho, (double)(global_timer_ns() - t_start) * 1e-6};
        a.out->log_rows = row + 1;
      }
    }
    if (chk) {
      bool ok;
      status = evaluate(1.0, ok);
      if (!ok) B200QP_DIE();
      if (status) break;
      if (last) {
        status = evaluate(10.0, ok);
        if (!ok) B200QP_DIE();
        if (!status) status = 7;
        break;
      }
    }
    if (tlp) { status = 8; break; }
    if (adp) {
      const double pr = s_prim / (s_nprim + 1e-10);
      const double du = s_dual / (s_ndual + 1e-10);
      double rho_new = rho * sqrt(pr / (du + 1e-10));
      rho_new = fmin(fmax(rho_new, RHO_MIN), RHO_MAX);
      if (rho_new > a.adapt_tol * rho || rho_new < rho / a.adapt_tol) {
        rho = rho_new;
        rho_updates++;
        // every rank re-derives the whole rho vector (the A' rows of its slice read all of it), ...
        for (int i = g0 + tid; i < g1; i += NT) {
          const double rv = rho_of_type(x.ctype_full[i], rho, a.rho_is_vec, a.rho_eq_factor);
          x.rho_full[i] = rv;
          x.rho_inv_full[i] = 1.0 / rv;
        }
        B200QP_BARRIER_OR_DIE();
        // ... refreshes w2 on its own rows and pushes it
        for (int i = m0 + tid; i < m1; i += NT) a.w2[i] = __ldcg(a.rho_v + i) * (a.z[i] - a.ax[i]) - a.y[i];
        __syncthreads();
        push_rows(x.w2f, w2full, mlo + m0, mlo + m1);
        B200QP_XBAR_OR_DIE(0, nullptr, 0u);
        if (!phase_R(false, true)) B200QP_DIE();
      }
    }

    // ---------------------------------------------------------------- PCG on K d = r (single-reduction recurrence)
    // a.p (full, gathered) holds u = M^-1 r: the R phase left u_0 there and <r_0, u_0> in rz
    bool have_d = false;
    if (r0max > 1e-150) {
      const double tol = a.cg_tol_frac * r0max;
      double gam = rz, gam_prev = rz, alpha_cg = 0.0;
      for (int j = 0; j < a.cg_max_iter; ++j) {
        // B (my rows): t = rho .* (A_g u), pushed to every rank
        for (int base = m0; base < m1; base += 2 * NGM) {
          const int rowA = base + grp_m, rowB = rowA + NGM;
          const bool okA = rowA < m1, okB = rowB < m1;
          double accA, accB;
          dot_pair<LGM, 3, 3>(a.A, a.p, rowA, okA, a.A, a.p, rowB, okB, lane_m, accA, accB);
          accA = group_sum<LGM>(accA);
          accB = group_sum<LGM>(accB);
          if (lane_m == 0) {
            if (okA) { const double v = __ldcg(a.rho_v + rowA) * accA; a.w[rowA] = v; push_one(x.wf, mlo + rowA, v); }
            if (okB) { const double v = __ldcg(a.rho_v + rowB) * accB; a.w[rowB] = v; push_one(x.wf, mlo + rowB, v); }
          }
        }
        B200QP_XBAR_OR_DIE(0, nullptr, 0u);
        mark(1);
        // C (my slice): w = K u = P u + sigma u + A' t ; <w, u> partial
        double v1[1] = {0.0};
        for (int base = s0; base < s1; base += NGN) {
          const int row = base + grp_n;
          const bool valid = row < s1;
          double accP, accA;
          dot_pair<LGN, 1, 4>(a.P, a.p, row, valid, a.At, wfull, row, valid, lane_n, accP, accA);
          const double acc = group_sum<LGN>(accP + accA);
          if (valid && lane_n == 0) {
            const double uj = __ldcg(a.p + row);
            const double wj = acc + sigma * uj;
            a.kp[row] = wj;
            v1[0] += uj * wj;
          }
        }
        double* bk = bank();
        block_reduce_store<1>(v1, 0u, bk, G, s_red);
        B200QP_XBAR_LOCAL_OR_DIE(1, bk, 0u);
        mark(2);
        const double delta = s_tot[0];
        const double bj = (j == 0) ? 0.0 : gam / gam_prev;
        const double pkp = (j == 0) ? delta : delta - bj * gam / alpha_cg;      // = <p, K p>
        if (!(pkp > 0.0)) { status = 9; break; }
        alpha_cg = gam / pkp;
        cg_total++;
        // V (my slice): p = u + b p ; s = w + b s ; d += a p ; r -= a s ; u = Minv r (pushed) ; <r, u> ; |r|inf
        double v2[2] = {0.0, 0.0};
        for (int i = s0 + tid; i < s1; i += NT) {
          const double ui = __ldcg(a.p + i), wi = a.kp[i];
          const double pi = (j == 0) ? ui : fma(bj, x.ps[i], ui);
          const double si = (j == 0) ? wi : fma(bj, x.ss[i], wi);
          x.ps[i] = pi; x.ss[i] = si;
          a.d[i] = (j == 0) ? alpha_cg * pi : fma(alpha_cg, pi, __ldcg(a.d + i));
          const double ri = fma(-alpha_cg, si, a.r[i]);
          a.r[i] = ri;
          const double un = a.Minv[i] * ri;
          a.p[i] = un;
          push_one(x.p, i, un);
          v2[0] += ri * un;
          v2[1] = nanmax(v2[1], fabs(ri));
        }
        have_d = true;
        bk = bank();
        block_reduce_store<2>(v2, 2u, bk, G, s_red);
        B200QP_XBAR_OR_DIE(2, bk, 2u);
        mark(3);
        gam_prev = gam;
        gam = s_tot[0];
        const double rmax = s_tot[1];
        if (!(rmax > tol) || j + 1 == a.cg_max_iter) break;
      }
      if (status) break;
    }
    if (have_d) {
      push_rows(x.d, a.d, s0, s1);               // the Z phase gathers from the whole step
    } else {
      for (int i = u0 + tid; i < u1; i += NT) a.d[i] = 0.0;
    }
    B200QP_XBAR_OR_DIE(0, nullptr, 0u);
    mark(6);

    // ---------------------------------------------------------------- Z phase (my rows)
    const int kn = k + 1;
    const bool last_n = (kn >= a.max_iter);
    const bool full_n = ((a.check_term > 0 && (kn % a.check_term) == 0) || last_n || tl_pending) ||
                        (!last_n && a.adaptive_rho && a.adapt_interval > 0 && (kn % a.adapt_interval) == 0);
    {
      double v[Z_K];
#pragma unroll
      for (int i = 0; i < Z_K; ++i) v[i] = 0.0;
      v[Z_VU] = NEG_INF; v[Z_VL] = NEG_INF;
      auto z_row = [&](int row, double ad) {
          const double ax_old = a.ax[row], z_old = a.z[row], y_old = a.y[row];
          const double rv = __ldcg(a.rho_v + row), ri = __ldcg(a.rho_inv + row), li = a.l[row], ui = a.u[row];
          const double zt = ax_old + ad;
          const double zr = alpha * zt + (1.0 - alpha) * z_old;
          const double zn = fmin(fmax(zr + ri * y_old, li), ui);
          const double dy = rv * (zr - zn);
          const double yn = y_old + dy;
          const double axn = fma(alpha, ad, ax_old);
          a.z[row] = zn; a.y[row] = yn; a.ax[row] = axn;
          const double w2n = rv * (zn - axn) - yn;
          a.w2[row] = w2n;
          push_one(x.w2f, mlo + row, w2n);
          if (full_n) {
            yfull[mlo + row] = yn;
            push_one(x.yf, mlo + row, yn);
            const double ei = Einv_t ? Einv_t[row] : 1.0;
            const double pr = axn - zn;
            v[Z_SPRIM] = nanmax(v[Z_SPRIM], fabs(pr));
            v[Z_SAX] = nanmax(v[Z_SAX], fabs(axn));
            v[Z_SZ] = nanmax(v[Z_SZ], fabs(zn));
            v[Z_TPRIM] = nanmax(v[Z_TPRIM], fabs(ei * pr));
            v[Z_TAX] = nanmax(v[Z_TAX], fabs(ei * axn));
            v[Z_TZ] = nanmax(v[Z_TZ], fabs(ei * zn));
            const bool lfin = li > -INF_THRESH, ufin = ui < INF_THRESH;
            double dyp = dy;
            if (!ufin && !lfin) dyp = 0.0;
            else if (!ufin) dyp = fmin(dy, 0.0);
            else if (!lfin) dyp = fmax(dy, 0.0);
            a.w[row] = dyp;
            push_one(x.wf, mlo + row, dyp);                  // projected delta_y, for the primal-infeasibility norm
            v[Z_NDY] = nanmax(v[Z_NDY], fabs((E_t ? E_t[row] : 1.0) * dyp));
            v[Z_LHS] += (ufin ? ui * fmax(dyp, 0.0) : 0.0) + (lfin ? li * fmin(dyp, 0.0) : 0.0);
            const double adx = ei * alpha * ad;
            if (ufin) v[Z_VU] = nanmax(v[Z_VU], adx);
            if (lfin) v[Z_VL] = nanmax(v[Z_VL], -adx);
            v[Z_SUP] += (ufin ? ui * fmax(yn, 0.0) : 0.0) + (lfin ? li * fmin(yn, 0.0) : 0.0);
          }
      };
      for (int base = m0; base < m1; base += 2 * NGM) {
        const int rowA = base + grp_m, rowB = rowA + NGM;
        const bool okA = rowA < m1, okB = rowB < m1;
        double adA, adB;
        dot_pair<LGM, 3, 3>(a.A, a.d, rowA, okA, a.A, a.d, rowB, okB, lane_m, adA, adB);
        adA = group_sum<LGM>(adA);
        adB = group_sum<LGM>(adB);
        if (lane_m == 0) {
          if (okA) z_row(rowA, adA);
          if (okB) z_row(rowB, adB);
        }
      }
      for (int i = u0 + tid; i < u1; i += NT) a.x[i] = fma(alpha, __ldcg(a.d + i), __ldcg(a.x + i));   // replicated
      if (full_n) {
        double* bk = bank();
        block_reduce_store<Z_K>(v, Z_MAXMASK, bk, G, s_red);
        B200QP_XBAR_OR_DIE(Z_K, bk, Z_MAXMASK);
        s_prim = s_tot[Z_SPRIM];
        s_nprim = fmax(s_tot[Z_SAX], s_tot[Z_SZ]);
        t_prim = s_tot[Z_TPRIM]; t_ax = s_tot[Z_TAX]; t_z = s_tot[Z_TZ];
        ndy = s_tot[Z_NDY]; lhs = s_tot[Z_LHS]; vu = s_tot[Z_VU]; vl = s_tot[Z_VL]; sup = s_tot[Z_SUP];
      } else {
        B200QP_XBAR_OR_DIE(0, nullptr, 0u);
      }
    }
    mark(5);
    k = kn;
  }

  if (cta == 0 && tid == 0) {
    KOut* o = a.out;
    for (int i = 0; i < 8; ++i) o->phase_ns[i] = t_phase[i];
    o->status = status; o->iter = k; o->rho_updates = rho_updates;
    o->cg_iters = cg_total; o->barriers = (long long)sync.epoch;
    o->xepoch = (long long)xs.xepoch;
    o->obj = obj; o->dual_obj = dobj; o->gap = gap; o->prim_res = prim_res; o->dual_res = dual_res;
    o->rho = rho;
    const double pr = s_prim / (s_nprim + 1e-10), du = s_dual / (s_ndual + 1e-10);
    o->rho_estimate = fmin(fmax(rho * sqrt(pr / (du + 1e-10)), RHO_MIN), RHO_MAX);
    o->rel_kkt = fmax(prim_res / (nprim + 1e-10), dual_res / (ndual + 1e-10));
  }
}

}  // namespace b200qp
