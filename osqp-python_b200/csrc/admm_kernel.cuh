// admm_kernel.cuh -- the persistent ADMM kernel: the whole osqp_solve() iteration loop on device.
//
// One cooperative launch runs every ADMM iteration of a solve (reference loop:
// src/osqppurepy/_osqp.py:1188-1306; upstream osqp_solve called at src/bindings.cpp.in:198):
//   R phase   n-rows : r = -(P x + q) + A'(rho(z - Ax) - y), Jacobi-preconditioned start p = M^-1 r;
//                      on check iterations also A'y and every norm of the termination test
//   PCG       C1 m-rows: w = rho .* (A p)          C2 n-rows: Kp = P p + sigma p + A' w, <p,Kp>
//             C3 n-elem: d += a p, r -= a Kp, <r,M^-1 r>, |r|inf      C4 n-elem: p = M^-1 r + b p
//   Z phase   m-rows : A d fused with relaxation, box projection, dual update, incremental Ax,
//                      next right-hand side rho(z - Ax) - y and all m-side norms
// Phases are separated by a grid barrier; reductions are deterministic (fixed order), so every
// CTA takes identical control decisions (CG exit, termination, adaptive rho) without host help.
//
// The linear system is solved in residual-correction form: K d = rhs - K x, xt = x + d, so the
// PCG starts from d = 0 and its tolerance is a fraction of the warm-start residual; the solver's
// hidden state is (x, z, y, rho) only (warm_start_test.py:49-52 pins that).
#pragma once
#include "device_common.cuh"

namespace b200qp {

struct LogRow { int iter; int pad; double obj, prim_res, dual_res, rho, t_ms; };

struct KOut {                 // written by CTA 0 at kernel exit, read back by the host
  int status; int iter; int rho_updates; int aborted;
  long long cg_iters; long long barriers;
  double obj, dual_obj, gap, prim_res, dual_res, rho_estimate, rho, rel_kkt;
  int log_rows; int pad;
  long long phase_ns[8];      // CTA-0 time per phase kind: R, C1, C2, C3, C4, Z, other (barrier waits included)
  long long xepoch;           // multi-GPU kernel: cross-GPU barriers completed (carried to the next launch)
};

struct KArgs {
  int n, m, G;
  CsrMat P, A, At;
  // n-vectors
  const double *q, *pdiag, *D, *Dinv;
  double *x, *d, *r, *p, *kp, *px, *Minv;
  // m-vectors
  const double *l, *u, *E, *Einv;
  const int* ctype;
  double *z, *y, *ax, *w2, *w, *rho_v, *rho_inv;
  // partitions (G+1 each), reduction banks [2][KMAX][G], control
  const int *part_n, *part_m;
  double* red;
  Ctrl* ctrl;
  KOut* out;
  LogRow* log; int log_cap;
  // settings
  double sigma, alpha, rho, eps_abs, eps_rel, eps_pinf, eps_dinf, c, adapt_tol, cg_tol_frac;
  double rho_eq_factor;       // rho on equality rows = rho_eq_factor * rho (reference: 1e3, _osqp.py:27)
  double nq_s, nq_t;          // |q|inf scaled, |Dinv q|inf (termination space)
  double time_limit_ns;
  int max_iter, check_term, adaptive_rho, adapt_interval, cg_max_iter, cg_precond;
  int unscale;                // termination in the unscaled space (scaling && !scaled_termination)
  int rho_is_vec, check_dualgap, verbose;
  // polish mode: run the same loop on the active-set problem (see solver: polish())
  int polish_mode;
};

// dot of one CSR row with a vector; LG lanes cooperate, partial sums returned per lane.
// The first two entries of every lane are fetched in straight-line code so that the index loads, then
// the gathers, of both entries (and of the other rows the caller interleaves) are in flight together.
template <int LG>
__device__ __forceinline__ double row_dot(const CsrMat& M, const double* xv, int row, bool valid, int lane) {
  double acc = 0.0;
  if (valid) {
    const int a = __ldg(M.ptr + row), b = __ldg(M.ptr + row + 1);
    const int k0 = a + lane, k1 = k0 + LG;
    const bool h0 = k0 < b, h1 = k1 < b;
    int j0 = 0, j1 = 0;
    double v0 = 0.0, v1 = 0.0;
    if (h0) { j0 = ld_stream(M.idx + k0); v0 = ld_stream(M.val + k0); }
    if (h1) { j1 = ld_stream(M.idx + k1); v1 = ld_stream(M.val + k1); }
    double x0 = 0.0, x1 = 0.0;
    if (h0) x0 = ld_vec(xv + j0);
    if (h1) x1 = ld_vec(xv + j1);
    acc = fma(v1, x1, v0 * x0);
#pragma unroll 4
    for (int k = k1 + LG; k < b; k += LG) acc = fma(ld_stream(M.val + k), ld_vec(xv + ld_stream(M.idx + k)), acc);
  }
  return acc;
}
// two independent row dots (different rows of one matrix, or one row of two matrices) with their loads
// interleaved.  The persistent kernel is latency-bound (ncu: 70 % long-scoreboard stalls, IPC 0.74), so every
// dependent round trip on the critical path counts: the first NP1 / NP2 entries of every lane are fetched in
// straight-line code -- all index + value loads first, then all gathers, then the FMAs -- and only rows longer
// than NP * LG entries enter the tail loop.  With 10 entries per row and 4 lanes, NP = 3 covers the whole row
// (the previous version prefetched 2 and paid two more serial index -> gather round trips per row pair).
// The row pointers come from the caller so that it can fetch those of the next pair one iteration ahead.
// Summation order per lane: entries in index order (unchanged).
template <int LG, int NP1, int NP2>
__device__ __forceinline__ void dot_pair_ptr(const CsrMat& M1, const double* x1, int a1, int b1, const CsrMat& M2,
                                             const double* x2, int a2, int b2, int lane, double& acc1, double& acc2) {
  int j1[NP1], j2[NP2];
  double v1[NP1], v2[NP2];
#pragma unroll
  for (int e = 0; e < NP1; ++e) {
    const int k = a1 + lane + e * LG;
    j1[e] = 0; v1[e] = 0.0;
    if (k < b1) { j1[e] = ld_stream(M1.idx + k); v1[e] = ld_stream(M1.val + k); }
  }
#pragma unroll
  for (int e = 0; e < NP2; ++e) {
    const int k = a2 + lane + e * LG;
    j2[e] = 0; v2[e] = 0.0;
    if (k < b2) { j2[e] = ld_stream(M2.idx + k); v2[e] = ld_stream(M2.val + k); }
  }
  double y1[NP1], y2[NP2];
#pragma unroll
  for (int e = 0; e < NP1; ++e) y1[e] = (a1 + lane + e * LG < b1) ? ld_vec(x1 + j1[e]) : 0.0;
#pragma unroll
  for (int e = 0; e < NP2; ++e) y2[e] = (a2 + lane + e * LG < b2) ? ld_vec(x2 + j2[e]) : 0.0;
  acc1 = v1[0] * y1[0];
#pragma unroll
  for (int e = 1; e < NP1; ++e) acc1 = fma(v1[e], y1[e], acc1);
  acc2 = v2[0] * y2[0];
#pragma unroll
  for (int e = 1; e < NP2; ++e) acc2 = fma(v2[e], y2[e], acc2);
#pragma unroll 4
  for (int k = a1 + lane + NP1 * LG; k < b1; k += LG) acc1 = fma(ld_stream(M1.val + k), ld_vec(x1 + ld_stream(M1.idx + k)), acc1);
#pragma unroll 4
  for (int k = a2 + lane + NP2 * LG; k < b2; k += LG) acc2 = fma(ld_stream(M2.val + k), ld_vec(x2 + ld_stream(M2.idx + k)), acc2);
}
// NR independent rows of ONE matrix against one vector, same scheme: all index + value loads, then all gathers,
// then the FMAs (entries in index order per lane); rows longer than NP * LG entries finish in a tail loop.
template <int LG, int NP, int NR>
__device__ __forceinline__ void dot_rows(const CsrMat& M, const double* xv, const int (&ra)[NR], const int (&rb)[NR], int lane,
                                         double (&acc)[NR]) {
  int j[NR][NP];
  double v[NR][NP], y[NR][NP];
#pragma unroll
  for (int e = 0; e < NP; ++e)
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int k = ra[r] + lane + e * LG;
      j[r][e] = 0; v[r][e] = 0.0;
      if (k < rb[r]) { j[r][e] = ld_stream(M.idx + k); v[r][e] = ld_stream(M.val + k); }
    }
#pragma unroll
  for (int e = 0; e < NP; ++e)
#pragma unroll
    for (int r = 0; r < NR; ++r) y[r][e] = (ra[r] + lane + e * LG < rb[r]) ? ld_vec(xv + j[r][e]) : 0.0;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    acc[r] = v[r][0] * y[r][0];
#pragma unroll
    for (int e = 1; e < NP; ++e) acc[r] = fma(v[r][e], y[r][e], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) {
#pragma unroll 4
    for (int k = ra[r] + lane + NP * LG; k < rb[r]; k += LG) acc[r] = fma(ld_stream(M.val + k), ld_vec(xv + ld_stream(M.idx + k)), acc[r]);
  }
}
template <int LG, int NP1 = 3, int NP2 = 3>
__device__ __forceinline__ void dot_pair(const CsrMat& M1, const double* x1, int row1, bool ok1, const CsrMat& M2,
                                         const double* x2, int row2, bool ok2, int lane, double& acc1, double& acc2) {
  int a1 = 0, b1 = 0, a2 = 0, b2 = 0;
  if (ok1) { a1 = __ldg(M1.ptr + row1); b1 = __ldg(M1.ptr + row1 + 1); }
  if (ok2) { a2 = __ldg(M2.ptr + row2); b2 = __ldg(M2.ptr + row2 + 1); }
  dot_pair_ptr<LG, NP1, NP2>(M1, x1, a1, b1, M2, x2, a2, b2, lane, acc1, acc2);
}
template <int LG>
__device__ __forceinline__ void row_dot2(const CsrMat& M, const double* x1, const double* x2, int row, bool valid,
                                         int lane, double& a1, double& a2) {
  a1 = 0.0; a2 = 0.0;
  if (valid) {
    const int a = __ldg(M.ptr + row), b = __ldg(M.ptr + row + 1);
    for (int k = a + lane; k < b; k += LG) {
      const double v = ld_stream(M.val + k);
      const int j = ld_stream(M.idx + k);
      a1 = fma(v, ld_vec(x1 + j), a1);
      a2 = fma(v, ld_vec(x2 + j), a2);
    }
  }
}
template <int LG>
__device__ __forceinline__ double row_dot_sq(const CsrMat& M, const double* xv, int row, bool valid, int lane) {
  double acc = 0.0;
  if (valid) {
    const int a = __ldg(M.ptr + row), b = __ldg(M.ptr + row + 1);
    for (int k = a + lane; k < b; k += LG) {
      const double v = ld_stream(M.val + k);
      acc = fma(v * v, ld_vec(xv + ld_stream(M.idx + k)), acc);
    }
  }
  return acc;
}

__device__ __forceinline__ double rho_of_type(int ct, double rho, int rho_is_vec, double eq_factor) {
  if (!rho_is_vec) return rho;
  return ct == 1 ? eq_factor * rho : (ct == -1 ? RHO_MIN : rho);
}

#define B200QP_BARRIER_OR_DIE() do { if (!sync.barrier()) { if (blockIdx.x == 0 && threadIdx.x == 0) a.out->aborted = 1; return; } } while (0)

// slots of the R phase
enum { R_RZ = 0, R_R0, R_SDUAL, R_SATY, R_SPX, R_TDUAL, R_TATY, R_TPX, R_XPX, R_QX, R_NDX, R_QDX, R_TL, R_K };
constexpr unsigned R_MAXMASK = (1u << R_R0) | (1u << R_SDUAL) | (1u << R_SATY) | (1u << R_SPX) | (1u << R_TDUAL) |
                               (1u << R_TATY) | (1u << R_TPX) | (1u << R_NDX) | (1u << R_TL);
// slots of the Z phase
enum { Z_SPRIM = 0, Z_SAX, Z_SZ, Z_TPRIM, Z_TAX, Z_TZ, Z_NDY, Z_LHS, Z_VU, Z_VL, Z_SUP, Z_K };
constexpr unsigned Z_MAXMASK = (1u << Z_SPRIM) | (1u << Z_SAX) | (1u << Z_SZ) | (1u << Z_TPRIM) | (1u << Z_TAX) |
                               (1u << Z_TZ) | (1u << Z_NDY) | (1u << Z_VU) | (1u << Z_VL);

template <int LGN, int LGM>
__global__ void __launch_bounds__(NT, 2) admm_kernel(const __grid_constant__ KArgs a) {
  __shared__ double s_red[KMAX * NW];
  __shared__ double s_tot[KMAX];
  __shared__ int s_ok;

  const int G = a.G, cta = blockIdx.x, tid = threadIdx.x;
  const int n0 = a.part_n[cta], n1 = a.part_n[cta + 1];
  const int m0 = a.part_m[cta], m1 = a.part_m[cta + 1];
  GridSync sync{a.ctrl, G, 0ULL, &s_ok};
  const double NEG_INF = __longlong_as_double(0xfff0000000000000LL);

  // per-row cooperative groups
  constexpr int NGN = NT / LGN, NGM = NT / LGM;
  const int lane_n = tid & (LGN - 1), grp_n = tid / LGN;
  const int lane_m = tid & (LGM - 1), grp_m = tid / LGM;

  const double alpha = a.alpha, sigma = a.sigma;
  const double cinv = 1.0 / a.c;
  const double c_eff = a.unscale ? a.c : 1.0, cinv_eff = a.unscale ? cinv : 1.0;
  const double* Dinv_t = a.unscale ? a.Dinv : nullptr;   // termination-space scalings
  const double* D_t = a.unscale ? a.D : nullptr;
  const double* Einv_t = a.unscale ? a.Einv : nullptr;
  const double* E_t = a.unscale ? a.E : nullptr;

  double rho = a.rho;
  int rho_updates = 0, status = 0, k = 0;
  long long cg_total = 0;
  const unsigned long long t_start = global_timer_ns();

  // latest residual information (identical in all threads)
  double obj = 0, dobj = 0, gap = 0, prim_res = 0, dual_res = 0, nprim = 0, ndual = 0;
  double s_prim = 0, s_nprim = 0, s_dual = 0, s_ndual = 0;
  double ndy = 0, lhs = 0, vu = NEG_INF, vl = NEG_INF, sup = 0;   // m-side, from the last Z phase
  double t_prim = 0, t_ax = 0, t_z = 0;
  double ndx = 0, qdx = 0;
  double rz = 0, r0max = 0;

  auto bank = [&]() -> double* { return a.red + (size_t)(sync.epoch & 1ULL) * KMAX * G; };

  // ------------------------------------------------------------------ init: ax = A x, w2
  for (int base = m0; base < m1; base += NGM) {
    const int row = base + grp_m;
    const bool valid = row < m1;
    const double acc = group_sum<LGM>(row_dot<LGM>(a.A, a.x, row, valid, lane_m));
    if (valid && lane_m == 0) {
      a.ax[row] = acc;
      a.w2[row] = a.rho_v[row] * (a.z[row] - acc) - a.y[row];
    }
  }
  B200QP_BARRIER_OR_DIE();

  // per-phase timing (CTA 0 only; a few globaltimer reads per phase)
  __shared__ long long t_phase[8];
  __shared__ unsigned long long t_mark;
  if (tid == 0) { for (int i = 0; i < 8; ++i) t_phase[i] = 0; t_mark = global_timer_ns(); }
  auto mark = [&](int id) {
    if (cta == 0 && tid == 0) { const unsigned long long now = global_timer_ns(); t_phase[id] += (long long)(now - t_mark); t_mark = now; }
  };

  // R phase body (full = with the residual quantities of the current iterate)
  auto phase_R = [&](bool full, bool lite) {
    // lite: px already stored, recompute only A'w2 (after a rho update); also refreshes Minv
    double v[R_K];
#pragma unroll
    for (int i = 0; i < R_K; ++i) v[i] = 0.0;
    for (int base = n0; base < n1; base += NGN) {
      const int row = base + grp_n;
      const bool valid = row < n1;
      double px = 0, atw = 0, aty = 0, msum = 0;
      if (lite) {
        atw = row_dot<LGN>(a.At, a.w2, row, valid, lane_n);
        if (a.cg_precond) msum = row_dot_sq<LGN>(a.At, a.rho_v, row, valid, lane_n);
        msum = group_sum<LGN>(msum);
      } else {
        if (full) {
          px = row_dot<LGN>(a.P, a.x, row, valid, lane_n);
          row_dot2<LGN>(a.At, a.w2, a.y, row, valid, lane_n, atw, aty);
        } else {
          dot_pair<LGN, 1, 4>(a.P, a.x, row, valid, a.At, a.w2, row, valid, lane_n, px, atw);
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
        v[R_RZ] += rj * pj;
        v[R_R0] = nanmax(v[R_R0], fabs(rj));
        if (full) {
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
      }
    }
    // the clock is sampled every iteration (upstream osqp_solve does): a check_termination = 0 solve must still stop
    if (tid == 0)
      v[R_TL] = ((double)(global_timer_ns() - t_start) > a.time_limit_ns) ? 1.0 : 0.0;
    double* bk = bank();
    block_reduce_store<R_K>(v, R_MAXMASK, bk, G, s_red);
    return bk;
  };

  bool time_up = false, tl_pending = false;
  auto load_R = [&](const double* bk, bool full) {
    grid_reduce_load<R_K>(bk, G, R_MAXMASK, s_tot);
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
  };

  // n-rows SpMV + inf-norm for the infeasibility certificates; returns the norm
  auto cert_norm = [&](const CsrMat& M, const double* vec, double scale, bool& ok) -> double {
    double v[1] = {0.0};
    for (int base = n0; base < n1; base += NGN) {
      const int row = base + grp_n;
      const bool valid = row < n1;
      const double acc = group_sum<LGN>(row_dot<LGN>(M, vec, row, valid, lane_n));
      if (valid && lane_n == 0) v[0] = nanmax(v[0], fabs((Dinv_t ? Dinv_t[row] : 1.0) * scale * acc));
    }
    double* bk = bank();
    block_reduce_store<1>(v, 1u, bk, G, s_red);
    ok = sync.barrier();
    if (!ok) return 0.0;
    grid_reduce_load<1>(bk, G, 1u, s_tot);
    return s_tot[0];
  };

  // termination test of the current iterate (reference: _osqp.py:998-1077); f = 10 -> "inaccurate"
  auto evaluate = [&](double f, bool& ok) -> int {
    ok = true;
    if (!(prim_res <= QP_INFTY) || !(dual_res <= QP_INFTY)) return 9;  // NON_CVX (also catches NaN)
    const double ea = f * a.eps_abs, er = f * a.eps_rel;
    const bool prim_ok = (a.m == 0) || (prim_res < ea + er * nprim);
    const bool dual_ok = dual_res < ea + er * ndual;
    const bool gap_ok = !a.check_dualgap || (fabs(gap) < ea + er * fmax(fabs(obj), fabs(dobj)));
    if (prim_ok && dual_ok && gap_ok) return f > 1.0 ? 2 : 1;
    if (a.polish_mode) return 0;
    if (!prim_ok) {
      const double eps = f * a.eps_pinf;
      if (ndy > eps && lhs < -eps * ndy) {
        const double nat = cert_norm(a.At, a.w, 1.0, ok);   // w holds the projected delta_y
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

  // ================================================================== main loop
  for (;;) {
    const bool last = (k >= a.max_iter);
    const bool tlp = tl_pending;      // the clock ran out during the previous iteration: evaluate this iterate, then stop
    const bool chk = k > 0 && ((a.check_term > 0 && (k % a.check_term) == 0) || last || tlp);
    const bool adp = k > 0 && !last && !tlp && a.adaptive_rho && a.adapt_interval > 0 && (k % a.adapt_interval) == 0;
    const bool full = chk || adp;

    {
      const double* bk = phase_R(full, false);
      B200QP_BARRIER_OR_DIE();
      load_R(bk, full);
      mark(0);
    }
    if (time_up) tl_pending = true;
    if (full && a.verbose && cta == 0 && tid == 0 && chk && !tlp) {
      const int row = (k / (a.check_term > 0 ? a.check_term : 1)) - 1;
      if (row >= 0 && row < a.log_cap) {
        a.log[row] = LogRow{k, 0, obj, prim_res, dual_res, rho, (double)(global_timer_ns() - t_start) * 1e-6};
        a.out->log_rows = row + 1;
      }
    }
    if (chk) {
      bool ok;
      status = evaluate(1.0, ok);
      if (!ok) { if (cta == 0 && tid == 0) a.out->aborted = 1; return; }
      if (status) break;
      if (last) {
        status = evaluate(10.0, ok);
        if (!ok) { if (cta == 0 && tid == 0) a.out->aborted = 1; return; }
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
        for (int i = m0 + tid; i < m1; i += NT) {
          const double rv = rho_of_type(a.ctype[i], rho, a.rho_is_vec, a.rho_eq_factor);
          a.rho_v[i] = rv;
          a.rho_inv[i] = 1.0 / rv;
          a.w2[i] = rv * (a.z[i] - a.ax[i]) - a.y[i];
        }
        B200QP_BARRIER_OR_DIE();
        const double* bk = phase_R(false, true);
        B200QP_BARRIER_OR_DIE();
        load_R(bk, false);
      }
    }

    // ---------------------------------------------------------------- PCG on K d = r
    bool have_d = false;
    if (r0max > 1e-150) {
      const double tol = a.cg_tol_frac * r0max;
      for (int j = 0; j < a.cg_max_iter; ++j) {
        // C1: w = rho .* (A p): C1_NR rows per lane group and iteration (row pointers of the next set are fetched
        // one iteration ahead); with 3 entries per lane that is 3 C1_NR index + value loads, then 3 C1_NR gathers,
        // in flight per lane
        {
          constexpr int C1_NR = 2;   // 3 rows measured slower (40.0 vs 36.7 ms per solve: register pressure at the 64-register cap)
          int na[C1_NR], nb[C1_NR];
#pragma unroll
          for (int r = 0; r < C1_NR; ++r) {
            const int row = m0 + grp_m + r * NGM;
            na[r] = 0; nb[r] = 0;
            if (row < m1) { na[r] = __ldg(a.A.ptr + row); nb[r] = __ldg(a.A.ptr + row + 1); }
          }
          for (int base = m0; base < m1; base += C1_NR * NGM) {
            int ca[C1_NR], cb[C1_NR];
#pragma unroll
            for (int r = 0; r < C1_NR; ++r) {
              ca[r] = na[r]; cb[r] = nb[r];
              const int nrow = base + grp_m + (r + C1_NR) * NGM;
              na[r] = 0; nb[r] = 0;
              if (nrow < m1) { na[r] = __ldg(a.A.ptr + nrow); nb[r] = __ldg(a.A.ptr + nrow + 1); }
            }
            double acc[C1_NR];
            dot_rows<LGM, 3, C1_NR>(a.A, a.p, ca, cb, lane_m, acc);
#pragma unroll
            for (int r = 0; r < C1_NR; ++r) {
              const double sacc = group_sum<LGM>(acc[r]);
              const int row = base + grp_m + r * NGM;
              if (lane_m == 0 && row < m1) a.w[row] = a.rho_v[row] * sacc;
            }
          }
        }
        B200QP_BARRIER_OR_DIE();
        mark(1);
        // C2: kp = P p + sigma p + A' w ; <p, kp>
        double v1[1] = {0.0};
        {
          int pp0 = 0, pp1 = 0, pt0 = 0, pt1 = 0;
          if (n0 + grp_n < n1) {
            const int row = n0 + grp_n;
            pp0 = __ldg(a.P.ptr + row); pp1 = __ldg(a.P.ptr + row + 1);
            pt0 = __ldg(a.At.ptr + row); pt1 = __ldg(a.At.ptr + row + 1);
          }
          for (int base = n0; base < n1; base += NGN) {
            const int row = base + grp_n;
            const bool valid = row < n1;
            const int cp0 = pp0, cp1 = pp1, ct0 = pt0, ct1 = pt1;
            const int nr = row + NGN;
            pp0 = pp1 = pt0 = pt1 = 0;
            if (nr < n1) {
              pp0 = __ldg(a.P.ptr + nr); pp1 = __ldg(a.P.ptr + nr + 1);
              pt0 = __ldg(a.At.ptr + nr); pt1 = __ldg(a.At.ptr + nr + 1);
            }
            double accP, accA;
            dot_pair_ptr<LGN, 1, 4>(a.P, a.p, cp0, cp1, a.At, a.w, ct0, ct1, lane_n, accP, accA);
            const double acc = group_sum<LGN>(accP + accA);
            if (valid && lane_n == 0) {
              const double pj = a.p[row];
              const double kpj = acc + sigma * pj;
              a.kp[row] = kpj;
              v1[0] += pj * kpj;
            }
          }
        }
        double* bk = bank();
        block_reduce_store<1>(v1, 0u, bk, G, s_red);
        B200QP_BARRIER_OR_DIE();
        grid_reduce_load<1>(bk, G, 0u, s_tot);
        mark(2);
        const double pkp = s_tot[0];
        if (!(pkp > 0.0)) { status = 9; break; }   // K not positive definite (or NaN): non-convex
        const double aj = rz / pkp;
        cg_total++;
        // C3: d += a p ; r -= a kp ; <r, Minv r> ; |r|inf
        double v2[2] = {0.0, 0.0};
        for (int i = n0 + tid; i < n1; i += NT) {
          const double pi = a.p[i];
          a.d[i] = (j == 0) ? aj * pi : fma(aj, pi, a.d[i]);
          const double ri = fma(-aj, a.kp[i], a.r[i]);
          a.r[i] = ri;
          v2[0] += ri * a.Minv[i] * ri;
          v2[1] = nanmax(v2[1], fabs(ri));
        }
        have_d = true;
        bk = bank();
        block_reduce_store<2>(v2, 2u, bk, G, s_red);
        B200QP_BARRIER_OR_DIE();
        grid_reduce_load<2>(bk, G, 2u, s_tot);
        mark(3);
        const double rz_new = s_tot[0], rmax = s_tot[1];
        if (!(rmax > tol) || j + 1 == a.cg_max_iter) break;
        // C4: p = Minv r + b p
        const double bj = rz_new / rz;
        rz = rz_new;
        for (int i = n0 + tid; i < n1; i += NT) a.p[i] = fma(bj, a.p[i], a.Minv[i] * a.r[i]);
        B200QP_BARRIER_OR_DIE();
        mark(4);
      }
      if (status) break;
    }
    if (!have_d) {
      for (int i = n0 + tid; i < n1; i += NT) a.d[i] = 0.0;
      B200QP_BARRIER_OR_DIE();
    }

    // ---------------------------------------------------------------- Z phase
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
          const double rv = a.rho_v[row], ri = a.rho_inv[row], li = a.l[row], ui = a.u[row];
          const double zt = ax_old + ad;                                   // z~ = A x~
          const double zr = alpha * zt + (1.0 - alpha) * z_old;            // _osqp.py:686-690
          const double zn = fmin(fmax(zr + ri * y_old, li), ui);           // _osqp.py:670-674
          const double dy = rv * (zr - zn);                                // _osqp.py:698-703
          const double yn = y_old + dy;
          const double axn = fma(alpha, ad, ax_old);
          a.z[row] = zn; a.y[row] = yn; a.ax[row] = axn;
          a.w2[row] = rv * (zn - axn) - yn;
          if (full_n) {
            const double ei = Einv_t ? Einv_t[row] : 1.0;
            const double pr = axn - zn;
            v[Z_SPRIM] = nanmax(v[Z_SPRIM], fabs(pr));
            v[Z_SAX] = nanmax(v[Z_SAX], fabs(axn));
            v[Z_SZ] = nanmax(v[Z_SZ], fabs(zn));
            v[Z_TPRIM] = nanmax(v[Z_TPRIM], fabs(ei * pr));
            v[Z_TAX] = nanmax(v[Z_TAX], fabs(ei * axn));
            v[Z_TZ] = nanmax(v[Z_TZ], fabs(ei * zn));
            // primal infeasibility (reference _osqp.py:796-820), delta_y projected on the polar of the
            // recession cone of [l,u] so infinite bounds never enter the support function
            const bool lfin = li > -INF_THRESH, ufin = ui < INF_THRESH;
            double dyp = dy;
            if (!ufin && !lfin) dyp = 0.0;
            else if (!ufin) dyp = fmin(dy, 0.0);
            else if (!lfin) dyp = fmax(dy, 0.0);
            a.w[row] = dyp;
            v[Z_NDY] = nanmax(v[Z_NDY], fabs((E_t ? E_t[row] : 1.0) * dyp));
            v[Z_LHS] += (ufin ? ui * fmax(dyp, 0.0) : 0.0) + (lfin ? li * fmin(dyp, 0.0) : 0.0);
            // dual infeasibility rows (reference _osqp.py:861-872): A dx must not point out of finite bounds
            const double adx = ei * alpha * ad;
            if (ufin) v[Z_VU] = nanmax(v[Z_VU], adx);
            if (lfin) v[Z_VL] = nanmax(v[Z_VL], -adx);
            // support function for the dual objective
            v[Z_SUP] += (ufin ? ui * fmax(yn, 0.0) : 0.0) + (lfin ? li * fmin(yn, 0.0) : 0.0);
          }
      };
      for (int base = m0; base < m1; base += 2 * NGM) {
        const int rowA = base + grp_m, rowB = rowA + NGM;
        const bool okA = rowA < m1, okB = rowB < m1;
        double adA, adB;
        dot_pair<LGM>(a.A, a.d, rowA, okA, a.A, a.d, rowB, okB, lane_m, adA, adB);
        adA = group_sum<LGM>(adA);
        adB = group_sum<LGM>(adB);
        if (lane_m == 0) {
          if (okA) z_row(rowA, adA);
          if (okB) z_row(rowB, adB);
        }
      }
      for (int i = n0 + tid; i < n1; i += NT) a.x[i] = fma(alpha, a.d[i], a.x[i]);  // _osqp.py:664-668
      if (full_n) {
        double* bk = bank();
        block_reduce_store<Z_K>(v, Z_MAXMASK, bk, G, s_red);
        B200QP_BARRIER_OR_DIE();
        grid_reduce_load<Z_K>(bk, G, Z_MAXMASK, s_tot);
        s_prim = s_tot[Z_SPRIM];
        s_nprim = fmax(s_tot[Z_SAX], s_tot[Z_SZ]);
        t_prim = s_tot[Z_TPRIM]; t_ax = s_tot[Z_TAX]; t_z = s_tot[Z_TZ];
        ndy = s_tot[Z_NDY]; lhs = s_tot[Z_LHS]; vu = s_tot[Z_VU]; vl = s_tot[Z_VL]; sup = s_tot[Z_SUP];
      } else {
        B200QP_BARRIER_OR_DIE();
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
    o->obj = obj; o->dual_obj = dobj; o->gap = gap; o->prim_res = prim_res; o->dual_res = dual_res;
    o->rho = rho;
    const double pr = s_prim / (s_nprim + 1e-10), du = s_dual / (s_ndual + 1e-10);
    o->rho_estimate = fmin(fmax(rho * sqrt(pr / (du + 1e-10)), RHO_MIN), RHO_MAX);
    o->rel_kkt = fmax(prim_res / (nprim + 1e-10), dual_res / (ndual + 1e-10));
  }
}

}  // namespace b200qp
