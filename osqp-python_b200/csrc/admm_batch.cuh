// admm_batch.cuh -- batched ADMM kernel: many independent QPs that share one (P, A), one CTA per QP.
//
// BASELINE.json configs[4] (8192 MPC QPs, n = 120, m = 240); SURVEY.md 8e "config 5: replicas only, no
// collective".  The reference solves such a batch by looping solver objects over a thread pool
// (src/osqp/nn/torch.py:128-164, 200-217), re-using one setup through update(q, l, u)
// (src/osqp/nn/torch.py:136-140).  Element b of the batch is exactly that QP: the shared scaled matrices of
// one setup(), its own q, l, u, its own ADMM state (x, z, y, rho).
//
// Same iteration as admm_kernel.cuh (reference loop src/osqppurepy/_osqp.py:1188-1306): residual-correction
// Jacobi-PCG on K = P + sigma I + A' rho A, relaxation, box projection, dual update, residuals, infeasibility
// certificates, adaptive rho.  What changes is where things live:
//   * every vector of the QP (8 n-vectors + 8 m-vectors) is in SHARED MEMORY for the whole solve; global
//     memory is touched once to load q, l, u (and the warm start) and once to store x, y and the info block;
//   * the shared matrices are read through L1 (__ldg): all CTAs of an SM walk the same 30 KB;
//   * phases are separated by __syncthreads, reductions are CTA-level and fixed-order (deterministic);
//   * CTAs pull QP indices from a global counter, so QPs that need 25 and 200 iterations balance out.
// Two inner solvers (template flag DENSE):
//   * DENSE = false: the engine's Jacobi-PCG, as in admm_kernel.cuh (any n, m whose vectors fit shared memory);
//   * DENSE = true : small QPs (n <= 128).  K is formed in shared memory and inverted IN REGISTERS once per rho
//     value (Gauss-Jordan; K is SPD so no pivoting, a non-positive pivot = NON_CVX) and the PCG is replaced by
//     one dense mat-vec d = K^-1 r per ADMM iteration -- 3 CTA barriers per iteration instead of ~30.  The
//     residual-correction form keeps the step exact to the accuracy of K^-1 relative to the correction d, not x.
//     This is the reference's "factor once, solve every iteration" scheme (src/osqppurepy/_osqp.py:281-312,
//     refactor on rho update :1597) with the factorisation held on chip.
#pragma once
#include "device_common.cuh"

namespace b200qp {

constexpr int NTB_PCG = 256;    // threads per CTA, PCG mode (4 CTAs per SM)
constexpr int NTB_DENSE = 512;  // threads per CTA, dense mode (1 CTA per SM: K^-1 takes most of the shared memory); a 1024-thread
                                // build exists for A/B runs (B200QP_BATCH_NT) and measured 23 % slower on config 5

struct BInfo {                  // per QP, written at exit (64 bytes)
  int status, iter, rho_updates, cg_iters;
  double obj, dual_obj, prim_res, dual_res, rho, rho_estimate;
};

struct BArgs {
  int n, m, batch;
  CsrMat P, A, At;                                   // shared, scaled (c D P D, E A D, D A' E)
  const double *pdiag, *D, *Dinv, *E, *Einv;         // shared
  const double *q, *l, *u;                           // [batch][n | m], scaled
  double *xs, *zs, *ys, *rho_b;                      // per-QP ADMM state, scaled space (kept between solves)
  double *sol_x, *sol_y, *cert_n, *cert_m;           // [batch][n | m], unscaled results
  double *kcache, *rcache;                           // dense mode, per CTA: K^-1 [n * ldk] of the last cold rho vector [m]
  const double *Pv, *Av, *Atv, *pdv;                 // per-element scaled matrix values [batch][nnz] (per_mat), else unused
  int per_mat;
  BInfo* info;
  int* next;                                         // work counter
  double sigma, alpha, eps_abs, eps_rel, eps_pinf, eps_dinf, c, adapt_tol, cg_tol_frac, rho_eq_factor, time_limit_ns;
  int max_iter, check_term, adaptive_rho, adapt_interval, cg_max_iter, cg_precond, unscale, rho_is_vec, check_dualgap;
  int warm;                                          // 1: start from (xs, ys), z = A x; 0: cold start
};

// CTA reduction of K slots (sum or NaN-propagating max per maxmask bit); result in s_tot[0..K)
template <int K, int NWB>
__device__ __forceinline__ void cta_reduce(double (&v)[K], unsigned maxmask, double* s_red, double* s_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double r = warp_combine(v[k], (maxmask >> k) & 1u);
    if (lane == 0) s_red[k * NWB + warp] = r;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    const bool mx = (maxmask >> threadIdx.x) & 1u;
    double acc = s_red[threadIdx.x * NWB];
#pragma unroll
    for (int w = 1; w < NWB; ++w) acc = combine(acc, s_red[threadIdx.x * NWB + w], mx);
    s_tot[threadIdx.x] = acc;
  }
  __syncthreads();
}

// dot of one shared-matrix row (through L1) with a shared-memory vector
template <int LG>
__device__ __forceinline__ double brow_dot(const CsrMat& M, const double* xs, int row, bool valid, int lane) {
  double acc = 0.0;
  if (valid) {
    const int a = __ldg(M.ptr + row), b = __ldg(M.ptr + row + 1);
    for (int k = a + lane; k < b; k += LG) acc = fma(__ldg(M.val + k), xs[__ldg(M.idx + k)], acc);
  }
  return acc;
}
template <int LG>
__device__ __forceinline__ void brow_dot2(const CsrMat& M, const double* x1, const double* x2, int row, bool valid,
                                          int lane, double& a1, double& a2) {
  a1 = 0.0; a2 = 0.0;
  if (valid) {
    const int a = __ldg(M.ptr + row), b = __ldg(M.ptr + row + 1);
    for (int k = a + lane; k < b; k += LG) {
      const double v = __ldg(M.val + k);
      const int j = __ldg(M.idx + k);
      a1 = fma(v, x1[j], a1);
      a2 = fma(v, x2[j], a2);
    }
  }
}
template <int LG>
__device__ __forceinline__ double brow_dot_sq(const CsrMat& M, const double* xs, int row, bool valid, int lane) {
  double acc = 0.0;
  if (valid) {
    const int a = __ldg(M.ptr + row), b = __ldg(M.ptr + row + 1);
    for (int k = a + lane; k < b; k += LG) { const double v = __ldg(M.val + k); acc = fma(v * v, xs[__ldg(M.idx + k)], acc); }
  }
  return acc;
}

__device__ __forceinline__ int btype_of(double li, double ui) {   // reference: _osqp.py:499-524
  return (ui - li < RHO_TOL) ? 1 : ((li < -INF_THRESH && ui > INF_THRESH) ? -1 : 0);
}
__device__ __forceinline__ double brho_of(int ct, double rho, int rho_is_vec, double eq_factor) {
  if (!rho_is_vec) return rho;
  return ct == 1 ? eq_factor * rho : (ct == -1 ? RHO_MIN : rho);
}

// reduction slots (same meaning as in admm_kernel.cuh)
enum { BR_RZ = 0, BR_R0, BR_SDUAL, BR_SATY, BR_SPX, BR_TDUAL, BR_TATY, BR_TPX, BR_XPX, BR_QX, BR_NDX, BR_QDX, BR_TL, BR_K };
constexpr unsigned BR_MAXMASK = (1u << BR_R0) | (1u << BR_SDUAL) | (1u << BR_SATY) | (1u << BR_SPX) | (1u << BR_TDUAL) |
                                (1u << BR_TATY) | (1u << BR_TPX) | (1u << BR_NDX) | (1u << BR_TL);
enum { BZ_SPRIM = 0, BZ_SAX, BZ_SZ, BZ_TPRIM, BZ_TAX, BZ_TZ, BZ_NDY, BZ_LHS, BZ_VU, BZ_VL, BZ_SUP, BZ_K };
constexpr unsigned BZ_MAXMASK = (1u << BZ_SPRIM) | (1u << BZ_SAX) | (1u << BZ_SZ) | (1u << BZ_TPRIM) | (1u << BZ_TAX) |
                                (1u << BZ_TZ) | (1u << BZ_NDY) | (1u << BZ_VU) | (1u << BZ_VL);

inline int batch_ldk(int n) { return n | 1; }                    // odd row stride of K^-1: column reads conflict-free
inline int batch_npad(int n) { return (n + 31) & ~31; }
inline size_t batch_smem_bytes(int n, int m, bool dense, int nt) {
  size_t d = (size_t)8 * n + (size_t)8 * m + KMAX * (nt / 32) + KMAX;
  if (dense) d += (size_t)n * batch_ldk(n) + 4 * 128 + (size_t)(nt / batch_npad(n)) * batch_npad(n);
  return d * sizeof(double);
}

constexpr int DENSE_NMAX = 128;
// One block of Gauss-Jordan steps on the register tile (see dense_build_inverse).  Thread (ty, tx) = (tid / TXW,
// tid % TXW), TXW = NT / 32, owns rows ty + 32 a (a < 4) and columns tx + TXW b (b < 128 / TXW).  Block BLK covers the
// TXW steps k = TXW BLK + 0 .. TXW - 1, for which the register index of the pivot row (AS) and of the pivot column
// (BLK) is a compile-time constant -- a run-time T[k >> 5][..] would push the whole tile into local memory.
template <int NT, int BLK>
__device__ __forceinline__ void gj_block(double (&T)[4][4096 / NT], double* line, int ty, int tx, int n, int& ok) {
  constexpr int TXW = NT / 32, NB = 4096 / NT;
  constexpr int K0 = TXW * BLK, AS = K0 >> 5;
#pragma unroll 1
  for (int kk = 0; kk < TXW; ++kk) {
    const int k = K0 + kk;
    if (k >= n) break;                               // uniform over the CTA
    double* rb = line + (k & 1) * 2 * DENSE_NMAX;    // double-buffered: one barrier per step
    double* cb = rb + DENSE_NMAX;
    const bool own_row = ty == (k & 31), own_col = tx == kk;
    if (own_row) {
#pragma unroll
      for (int bb = 0; bb < NB; ++bb) rb[tx + TXW * bb] = T[AS][bb];
    }
    if (own_col) {
#pragma unroll
      for (int aa = 0; aa < 4; ++aa) cb[ty + 32 * aa] = T[aa][BLK];
    }
    __syncthreads();
    const double akk = rb[k];
    if (!(akk > 0.0)) ok = 0;
    const double piv = 1.0 / akk;
    double rj[NB], ci[4];
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) rj[bb] = rb[tx + TXW * bb] * piv;
#pragma unroll
    for (int aa = 0; aa < 4; ++aa) ci[aa] = cb[ty + 32 * aa];
#pragma unroll
    for (int aa = 0; aa < 4; ++aa)
#pragma unroll
      for (int bb = 0; bb < NB; ++bb) T[aa][bb] = fma(-ci[aa], rj[bb], T[aa][bb]);
    if (own_row) {                                   // row k: a_kj / a_kk
#pragma unroll
      for (int bb = 0; bb < NB; ++bb) T[AS][bb] = rj[bb];
    }
    if (own_col) {                                   // column k: -a_ik / a_kk, and 1 / a_kk on the diagonal
#pragma unroll
      for (int aa = 0; aa < 4; ++aa) T[aa][BLK] = (aa == AS && own_row) ? piv : -ci[aa] * piv;
    }
  }
}
template <int NT, int BLK>
__device__ __forceinline__ void gj_blocks(double (&T)[4][4096 / NT], double* line, int ty, int tx, int n, int& ok) {
  gj_block<NT, BLK>(T, line, ty, tx, n, ok);
  if constexpr (BLK + 1 < 4096 / NT) gj_blocks<NT, BLK + 1>(T, line, ty, tx, n, ok);
}

// ---- dense mode: K = P + sigma I + A' rho A formed in shared memory (thread j owns row j: no atomics), then
// inverted by Gauss-Jordan elimination without pivoting (K is SPD; a non-positive pivot means it is not) with
// the whole matrix held in REGISTERS: with TXW = NT / 32, thread (ty, tx) = (tid / TXW, tid % TXW) owns rows
// ty + 32 a (a < 4) and columns tx + TXW b (b < 128 / TXW) of a 128 x 128 frame, so one elimination step is 4096 / NT
// FMAs per thread on registers plus a few broadcast reads of the pivot row / column from a double-buffered
// shared-memory line (one barrier per step).  Row k and column k of the step are then overwritten with their
// exact values.  Returns 0 if a pivot was not positive.  Requires n <= 128 and NT = 512 or 1024.
template <int NT>
__device__ __noinline__ int dense_build_inverse(const CsrMat P, const CsrMat A, const CsrMat At, int off_rho, double sigma,
                                                int n, int LDK, int off_K, int off_line) {
  static_assert(NT == 512 || NT == 1024, "register tiling assumes 32 x 16 or 32 x 32 threads");
  constexpr int TXW = NT / 32, NB = 4096 / NT;
  // offsets into the kernel's dynamic shared memory, not pointers: a pointer argument of a non-inlined function
  // is generic and every access would go through the L1TEX generic path instead of LDS / STS
  extern __shared__ double sm_dense[];
  const double* rho_v = sm_dense + off_rho;
  double* Kinv = sm_dense + off_K;
  double* line = sm_dense + off_line;
  const int tid = threadIdx.x;
  for (int e = tid; e < n * LDK; e += NT) Kinv[e] = 0.0;
  __syncthreads();
  if (tid < n) {
    double* Kj = Kinv + tid * LDK;
    for (int k = __ldg(P.ptr + tid); k < __ldg(P.ptr + tid + 1); ++k) Kj[__ldg(P.idx + k)] += __ldg(P.val + k);
    Kj[tid] += sigma;
    for (int t = __ldg(At.ptr + tid); t < __ldg(At.ptr + tid + 1); ++t) {
      const int i = __ldg(At.idx + t);
      const double f = rho_v[i] * __ldg(At.val + t);
      for (int k = __ldg(A.ptr + i); k < __ldg(A.ptr + i + 1); ++k) {
        const int c = __ldg(A.idx + k);
        Kj[c] = fma(f, __ldg(A.val + k), Kj[c]);
      }
    }
  }
  __syncthreads();
  const int ty = tid / TXW, tx = tid % TXW;
  double T[4][NB];
#pragma unroll
  for (int aa = 0; aa < 4; ++aa)
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) {
      const int i = ty + 32 * aa, j = tx + TXW * bb;
      T[aa][bb] = (i < n && j < n) ? Kinv[i * LDK + j] : 0.0;
    }
  // line: [2][2][DENSE_NMAX] = {buffer}{row, column}
  int ok = 1;
  gj_blocks<NT, 0>(T, line, ty, tx, n, ok);
  __syncthreads();
#pragma unroll
  for (int aa = 0; aa < 4; ++aa)
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) {
      const int i = ty + 32 * aa, j = tx + TXW * bb;
      if (i < n && j < n) Kinv[i * LDK + j] = T[aa][bb];
    }
  __syncthreads();
  return ok;
}

template <int LGN, int LGM, int NTB, bool DENSE>
__global__ void __launch_bounds__(NTB, DENSE ? 1 : 4) admm_batch_kernel(const __grid_constant__ BArgs a) {
  constexpr int NWB = NTB / 32;
  extern __shared__ double sm[];
  __shared__ int s_qp;
  __shared__ int s_tl;
  const int n = a.n, m = a.m, tid = threadIdx.x;
  double* const q = sm;
  double* const x = q + n;
  double* const d = x + n;
  double* const r = d + n;
  double* const p = r + n;
  double* const kp = p + n;
  double* const px = kp + n;
  double* const Minv = px + n;
  double* const l = Minv + n;
  double* const u = l + m;
  double* const z = u + m;
  double* const y = z + m;
  double* const ax = y + m;
  double* const w2 = ax + m;
  double* const w = w2 + m;
  double* const rho_v = w + m;
  double* const s_red = rho_v + m;
  double* const s_tot = s_red + KMAX * NWB;
  // dense mode: K^-1 (n x LDK), pivot lines of the elimination, partial sums of the mat-vec
  const int LDK = n | 1, npad = (n + 31) & ~31, PARTS = DENSE ? NTB / npad : 1;
  double* const Kinv = s_tot + KMAX;
  double* const rowk = Kinv + (DENSE ? n * LDK : 0);          // [2][2][128] pivot row / column lines of the elimination
  double* const partbuf = rowk + (DENSE ? 4 * DENSE_NMAX : 0);
  const int tj = tid % npad, ti0 = tid / npad;

  constexpr int NGN = NTB / LGN, NGM = NTB / LGM;
  const int lane_n = tid & (LGN - 1), grp_n = tid / LGN;
  const int lane_m = tid & (LGM - 1), grp_m = tid / LGM;
  const double NEG_INF = __longlong_as_double(0xfff0000000000000LL);
  const double NANV = __longlong_as_double(0x7ff8000000000000LL);
  const double alpha = a.alpha, sigma = a.sigma;
  const double cinv = 1.0 / a.c;
  const double c_eff = a.unscale ? a.c : 1.0, cinv_eff = a.unscale ? cinv : 1.0;
  const double* Dinv_t = a.unscale ? a.Dinv : nullptr;
  const double* D_t = a.unscale ? a.D : nullptr;
  const double* Einv_t = a.unscale ? a.Einv : nullptr;
  const double* E_t = a.unscale ? a.E : nullptr;

  int have_cache = 0;                      // dense mode: this CTA has stored an inverse in a.kcache
  for (;;) {
    __syncthreads();                       // the previous QP is completely stored
    if (tid == 0) s_qp = atomicAdd(a.next, 1);
    __syncthreads();
    const int qp = s_qp;
    if (qp >= a.batch) break;
    const unsigned long long t_start = global_timer_ns();
    const size_t on = (size_t)qp * n, om = (size_t)qp * m;
    // matrices of this element: shared pattern, shared or per-element values
    CsrMat P = a.P, A = a.A, At = a.At;
    const double* pdiag = a.pdiag;
    if (a.per_mat) {
      P.val = a.Pv + (size_t)qp * P.nnz; A.val = a.Av + (size_t)qp * A.nnz; At.val = a.Atv + (size_t)qp * At.nnz;
      pdiag = a.pdv + on;
    }

    // ---------------------------------------------------------------- load the QP
    double rho = a.rho_b[qp];
    double nq_s = 0.0, nq_t = 0.0;
    {
      double v[2] = {0.0, 0.0};
      for (int i = tid; i < n; i += NTB) {
        const double qi = a.q[on + i];
        q[i] = qi;
        x[i] = a.warm ? a.xs[on + i] : 0.0;
        d[i] = 0.0;
        v[0] = nanmax(v[0], fabs(qi));
        v[1] = nanmax(v[1], fabs((Dinv_t ? __ldg(Dinv_t + i) : 1.0) * qi));
      }
      for (int i = tid; i < m; i += NTB) {
        const double li = a.l[om + i], ui = a.u[om + i];
        l[i] = li; u[i] = ui;
        y[i] = a.warm ? a.ys[om + i] : 0.0;
        z[i] = a.warm ? a.zs[om + i] : 0.0;
        rho_v[i] = brho_of(btype_of(li, ui), rho, a.rho_is_vec, a.rho_eq_factor);
        w[i] = 0.0;
      }
      cta_reduce<2, NWB>(v, 3u, s_red, s_tot);
      nq_s = s_tot[0];
      nq_t = a.unscale ? s_tot[1] : s_tot[0];
    }
    // Jacobi diagonal of K and ax = A x, w2 = rho (z - ax) - y
    for (int base = 0; base < n; base += NGN) {
      const int row = base + grp_n;
      const bool valid = row < n;
      const bool jac = a.cg_precond && !DENSE;
      double ms = jac ? brow_dot_sq<LGN>(At, rho_v, row, valid, lane_n) : 0.0;
      ms = group_sum<LGN>(ms);
      if (valid && lane_n == 0) Minv[row] = jac ? 1.0 / (__ldg(pdiag + row) + sigma + ms) : 1.0;
    }
    for (int base = 0; base < m; base += NGM) {
      const int row = base + grp_m;
      const bool valid = row < m;
      const double acc = group_sum<LGM>(a.warm ? brow_dot<LGM>(A, x, row, valid, lane_m) : 0.0);
      if (valid && lane_m == 0) {
        ax[row] = acc;
        w2[row] = rho_v[row] * (z[row] - acc) - y[row];
      }
    }
    __syncthreads();

    int rho_updates = 0, status = 0, k = 0, cg_total = 0;
    double obj = 0, dobj = 0, gap = 0, prim_res = 0, dual_res = 0, nprim = 0, ndual = 0;
    double s_prim = 0, s_nprim = 0, s_dual = 0, s_ndual = 0;
    double ndy = 0, lhs = 0, vu = NEG_INF, vl = NEG_INF, sup = 0;
    double t_prim = 0, t_ax = 0, t_z = 0;
    double ndx = 0, qdx = 0;
    double rz = 0, r0max = 0;
    bool time_up = false, tl_stop = false, chk_next = false, adp_next = false;
    int cnt_chk = 0, cnt_adp = 0;

    // ---- dense mode: K = P + sigma I + A' rho A formed row by row (thread j owns row j: no atomics), inverted
    // in place by Gauss-Jordan elimination without pivoting (K is SPD; a non-positive pivot means it is not)
    auto build_inverse = [&]() -> bool {
      if constexpr (DENSE) return dense_build_inverse<NTB>(P, A, At, (int)(rho_v - sm), sigma, n, LDK, (int)(Kinv - sm), (int)(rowk - sm)) != 0;
      else return true;
    };
    // d = K^-1 r  (K^-1 symmetric: thread (part, j) walks column j = row j with unit-stride lanes)
    auto dense_apply = [&]() {
      if (tj < n && ti0 < PARTS) {
        double acc0 = 0.0, acc1 = 0.0;
        int t = ti0;
        for (; t + PARTS < n; t += 2 * PARTS) {
          acc0 = fma(Kinv[t * LDK + tj], r[t], acc0);
          acc1 = fma(Kinv[(t + PARTS) * LDK + tj], r[t + PARTS], acc1);
        }
        if (t < n) acc0 = fma(Kinv[t * LDK + tj], r[t], acc0);
        partbuf[ti0 * npad + tj] = acc0 + acc1;
      }
      __syncthreads();
      if (tid < n) {
        double sacc = partbuf[tid];
        for (int pp = 1; pp < PARTS; ++pp) sacc += partbuf[pp * npad + tid];
        d[tid] = sacc;
      }
      __syncthreads();
    };
    if (DENSE) {
      // K^-1 depends on the rho vector only.  QPs of one batch usually start from the same one (same constraint
      // types, rho = settings.rho), so every CTA keeps the first inverse it computes in global memory (L2) and
      // re-loads it -- 0.1 MB from L2 instead of n Gauss-Jordan steps -- whenever the rho vector matches.
      double* kc = a.kcache + (size_t)blockIdx.x * n * LDK;
      double* rc = a.rcache + (size_t)blockIdx.x * m;
      int same = have_cache && !a.per_mat;
      if (have_cache)
        for (int i = tid; i < m; i += NTB) same = same && (rc[i] == rho_v[i]);
      same = __syncthreads_and(same);
      if (same) {
        for (int e = tid; e < n * LDK; e += NTB) Kinv[e] = kc[e];
        __syncthreads();
      } else {
        if (!build_inverse()) status = 9;
        if (!have_cache && !a.per_mat && status == 0) {
          for (int e = tid; e < n * LDK; e += NTB) kc[e] = Kinv[e];
          for (int i = tid; i < m; i += NTB) rc[i] = rho_v[i];
          have_cache = 1;
          __syncthreads();
        }
      }
    }

    // R phase: r = -(P x + q) + A' w2, p = Minv r; full = also the dual-side residual quantities;
    // lite = after a rho update (px kept, Jacobi diagonal refreshed)
    auto phase_R = [&](bool full, bool lite) {
      double v[BR_K];
#pragma unroll
      for (int i = 0; i < BR_K; ++i) v[i] = 0.0;
      for (int base = 0; base < n; base += NGN) {
        const int row = base + grp_n;
        const bool valid = row < n;
        double pxv = 0, atw = 0, aty = 0, msum = 0;
        if (lite) {
          atw = brow_dot<LGN>(At, w2, row, valid, lane_n);
          if (a.cg_precond && !DENSE) msum = brow_dot_sq<LGN>(At, rho_v, row, valid, lane_n);
          msum = group_sum<LGN>(msum);
        } else {
          pxv = brow_dot<LGN>(P, x, row, valid, lane_n);
          if (full) brow_dot2<LGN>(At, w2, y, row, valid, lane_n, atw, aty);
          else atw = brow_dot<LGN>(At, w2, row, valid, lane_n);
          pxv = group_sum<LGN>(pxv);
          if (full) aty = group_sum<LGN>(aty);
        }
        atw = group_sum<LGN>(atw);
        if (valid && lane_n == 0) {
          const double qj = q[row];
          if (lite) {
            pxv = px[row];
            if (a.cg_precond && !DENSE) Minv[row] = 1.0 / (__ldg(pdiag + row) + sigma + msum);
          }
          const double rj = -(pxv + qj) + atw;
          const double pj = Minv[row] * rj;
          r[row] = rj;
          p[row] = pj;
          v[BR_RZ] += rj * pj;
          v[BR_R0] = nanmax(v[BR_R0], fabs(rj));
          if (!lite) px[row] = pxv;
          if (full) {
            const double xj = x[row], dj = alpha * d[row];
            const double dres = pxv + qj + aty;
            const double di = Dinv_t ? __ldg(Dinv_t + row) : 1.0;
            v[BR_SDUAL] = nanmax(v[BR_SDUAL], fabs(dres));
            v[BR_SATY] = nanmax(v[BR_SATY], fabs(aty));
            v[BR_SPX] = nanmax(v[BR_SPX], fabs(pxv));
            v[BR_TDUAL] = nanmax(v[BR_TDUAL], fabs(di * dres));
            v[BR_TATY] = nanmax(v[BR_TATY], fabs(di * aty));
            v[BR_TPX] = nanmax(v[BR_TPX], fabs(di * pxv));
            v[BR_XPX] += xj * pxv;
            v[BR_QX] += qj * xj;
            v[BR_NDX] = nanmax(v[BR_NDX], fabs((D_t ? __ldg(D_t + row) : 1.0) * dj));
            v[BR_QDX] += qj * dj;
          }
        }
      }
      // the clock is sampled every iteration (upstream osqp_solve does; check_termination = 0 must still stop)
      if (DENSE && !full) {          // the mat-vec needs neither <r, p> nor |r|inf: no reduction on plain iterations
        if (tid == 0) s_tl = ((double)(global_timer_ns() - t_start) > a.time_limit_ns) ? 1 : 0;
        __syncthreads();
        time_up = s_tl != 0;
        r0max = 1.0;
        return;
      }
      if (tid == 0) v[BR_TL] = ((double)(global_timer_ns() - t_start) > a.time_limit_ns) ? 1.0 : 0.0;
      cta_reduce<BR_K, NWB>(v, BR_MAXMASK, s_red, s_tot);
      rz = s_tot[BR_RZ];
      r0max = s_tot[BR_R0];
      time_up = s_tot[BR_TL] > 0.5;
      if (full) {
        s_dual = s_tot[BR_SDUAL];
        s_ndual = fmax(fmax(s_tot[BR_SATY], s_tot[BR_SPX]), nq_s);
        dual_res = cinv_eff * s_tot[BR_TDUAL];
        ndual = cinv_eff * fmax(fmax(s_tot[BR_TATY], s_tot[BR_TPX]), nq_t);
        obj = (0.5 * s_tot[BR_XPX] + s_tot[BR_QX]) * cinv;
        dobj = (-0.5 * s_tot[BR_XPX] - sup) * cinv;
        gap = obj - dobj;
        ndx = s_tot[BR_NDX];
        qdx = s_tot[BR_QDX];
        prim_res = t_prim;
        nprim = fmax(t_ax, t_z);
      }
    };

    // |Dinv * scale * (M vec)|inf over the n rows (infeasibility certificates)
    auto cert_norm = [&](const CsrMat& M, const double* vec, double scale) -> double {
      double v[1] = {0.0};
      for (int base = 0; base < n; base += NGN) {
        const int row = base + grp_n;
        const bool valid = row < n;
        const double acc = group_sum<LGN>(brow_dot<LGN>(M, vec, row, valid, lane_n));
        if (valid && lane_n == 0) v[0] = nanmax(v[0], fabs((Dinv_t ? __ldg(Dinv_t + row) : 1.0) * scale * acc));
      }
      cta_reduce<1, NWB>(v, 1u, s_red, s_tot);
      return s_tot[0];
    };

    // termination test (reference: _osqp.py:998-1077); f = 10 -> "inaccurate"
    auto evaluate = [&](double f) -> int {
      if (!(prim_res <= QP_INFTY) || !(dual_res <= QP_INFTY)) return 9;
      const double ea = f * a.eps_abs, er = f * a.eps_rel;
      const bool prim_ok = (m == 0) || (prim_res < ea + er * nprim);
      const bool dual_ok = dual_res < ea + er * ndual;
      const bool gap_ok = !a.check_dualgap || (fabs(gap) < ea + er * fmax(fabs(obj), fabs(dobj)));
      if (prim_ok && dual_ok && gap_ok) return f > 1.0 ? 2 : 1;
      if (!prim_ok) {
        const double eps = f * a.eps_pinf;
        if (ndy > eps && lhs < -eps * ndy) {
          const double nat = cert_norm(At, w, 1.0);
          if (nat < eps * ndy) return f > 1.0 ? 4 : 3;
        }
      }
      if (!dual_ok) {
        const double eps = f * a.eps_dinf;
        if (ndx > eps && qdx < -c_eff * eps * ndx && vu <= eps * ndx && vl <= eps * ndx) {
          const double npd = cert_norm(P, d, alpha);
          if (npd < c_eff * eps * ndx) return f > 1.0 ? 6 : 5;
        }
      }
      return 0;
    };

    // ================================================================ main loop
    for (; status == 0;) {
      const bool last = (k >= a.max_iter);
      const bool chk = chk_next, adp = adp_next;     // cadence of iteration k, counted in the previous Z phase
      const bool full = chk || adp;

      phase_R(full, false);
      if (chk) {
        status = evaluate(1.0);
        if (status) break;
        if (last) {
          status = evaluate(10.0);
          if (!status) status = 7;
          break;
        }
        if (tl_stop) { status = 8; break; }   // the clock ran out during the previous iteration
      }
      if (adp) {   // reference: _osqp.py:880-930 (scaled quantities)
        const double pr = s_prim / (s_nprim + 1e-10);
        const double du = s_dual / (s_ndual + 1e-10);
        double rho_new = rho * sqrt(pr / (du + 1e-10));
        rho_new = fmin(fmax(rho_new, RHO_MIN), RHO_MAX);
        if (rho_new > a.adapt_tol * rho || rho_new < rho / a.adapt_tol) {
          rho = rho_new;
          rho_updates++;
          for (int i = tid; i < m; i += NTB) {
            const double rv = brho_of(btype_of(l[i], u[i]), rho, a.rho_is_vec, a.rho_eq_factor);
            rho_v[i] = rv;
            w2[i] = rv * (z[i] - ax[i]) - y[i];
          }
          __syncthreads();
          if (DENSE && !build_inverse()) { status = 9; break; }
          phase_R(false, true);
        }
      }

      // -------------------------------------------------------------- PCG on K d = r
      bool have_d = false;
      if (DENSE) {
        if (r0max > 1e-150) { dense_apply(); have_d = true; }
      } else if (r0max > 1e-150) {
        const double tol = a.cg_tol_frac * r0max;
        for (int j = 0; j < a.cg_max_iter; ++j) {
          // C1: w = rho .* (A p)
          for (int base = 0; base < m; base += NGM) {
            const int row = base + grp_m;
            const bool valid = row < m;
            const double acc = group_sum<LGM>(brow_dot<LGM>(A, p, row, valid, lane_m));
            if (valid && lane_m == 0) w[row] = rho_v[row] * acc;
          }
          __syncthreads();
          // C2: kp = P p + sigma p + A' w ; <p, kp>
          double v1[1] = {0.0};
          for (int base = 0; base < n; base += NGN) {
            const int row = base + grp_n;
            const bool valid = row < n;
            const double acc = group_sum<LGN>(brow_dot<LGN>(P, p, row, valid, lane_n) + brow_dot<LGN>(At, w, row, valid, lane_n));
            if (valid && lane_n == 0) {
              const double pj = p[row];
              const double kpj = acc + sigma * pj;
              kp[row] = kpj;
              v1[0] += pj * kpj;
            }
          }
          cta_reduce<1, NWB>(v1, 0u, s_red, s_tot);
          const double pkp = s_tot[0];
          if (!(pkp > 0.0)) { status = 9; break; }   // K not positive definite (or NaN): non-convex
          const double aj = rz / pkp;
          cg_total++;
          // C3: d += a p ; r -= a kp ; <r, Minv r> ; |r|inf
          double v2[2] = {0.0, 0.0};
          for (int i = tid; i < n; i += NTB) {
            const double pi = p[i];
            d[i] = (j == 0) ? aj * pi : fma(aj, pi, d[i]);
            const double ri = fma(-aj, kp[i], r[i]);
            r[i] = ri;
            v2[0] += ri * Minv[i] * ri;
            v2[1] = nanmax(v2[1], fabs(ri));
          }
          have_d = true;
          cta_reduce<2, NWB>(v2, 2u, s_red, s_tot);
          const double rz_new = s_tot[0], rmax = s_tot[1];
          if (!(rmax > tol) || j + 1 == a.cg_max_iter) break;
          // C4: p = Minv r + b p
          const double bj = rz_new / rz;
          rz = rz_new;
          for (int i = tid; i < n; i += NTB) p[i] = fma(bj, p[i], Minv[i] * r[i]);
          __syncthreads();
        }
        if (status) break;
      }
      if (!have_d) {
        for (int i = tid; i < n; i += NTB) d[i] = 0.0;
        __syncthreads();
      }

      // -------------------------------------------------------------- Z phase
      // check every check_term-th, adapt every adapt_interval-th iteration (counters: no integer division per iteration)
      const int kn = k + 1;
      const bool last_n = (kn >= a.max_iter);
      bool hit_chk = false, hit_adp = false;
      if (++cnt_chk == a.check_term) { cnt_chk = 0; hit_chk = true; }
      if (++cnt_adp == a.adapt_interval) { cnt_adp = 0; hit_adp = true; }
      tl_stop = time_up;
      chk_next = hit_chk || last_n || tl_stop;
      adp_next = !last_n && !tl_stop && a.adaptive_rho && hit_adp;
      const bool full_n = chk_next || adp_next;
      {
        double v[BZ_K];
#pragma unroll
        for (int i = 0; i < BZ_K; ++i) v[i] = 0.0;
        v[BZ_VU] = NEG_INF; v[BZ_VL] = NEG_INF;
        for (int base = 0; base < m; base += NGM) {
          const int row = base + grp_m;
          const bool valid = row < m;
          const double ad = group_sum<LGM>(brow_dot<LGM>(A, d, row, valid, lane_m));
          if (valid && lane_m == 0) {
            const double ax_old = ax[row], z_old = z[row], y_old = y[row];
            const double rv = rho_v[row], ri = 1.0 / rv, li = l[row], ui = u[row];
            const double zt = ax_old + ad;                                   // z~ = A x~
            const double zr = alpha * zt + (1.0 - alpha) * z_old;            // _osqp.py:686-690
            const double zn = fmin(fmax(zr + ri * y_old, li), ui);           // _osqp.py:670-674
            const double dy = rv * (zr - zn);                                // _osqp.py:698-703
            const double yn = y_old + dy;
            const double axn = fma(alpha, ad, ax_old);
            z[row] = zn; y[row] = yn; ax[row] = axn;
            w2[row] = rv * (zn - axn) - yn;
            if (full_n) {
              const double ei = Einv_t ? __ldg(Einv_t + row) : 1.0;
              const double pr = axn - zn;
              v[BZ_SPRIM] = nanmax(v[BZ_SPRIM], fabs(pr));
              v[BZ_SAX] = nanmax(v[BZ_SAX], fabs(axn));
              v[BZ_SZ] = nanmax(v[BZ_SZ], fabs(zn));
              v[BZ_TPRIM] = nanmax(v[BZ_TPRIM], fabs(ei * pr));
              v[BZ_TAX] = nanmax(v[BZ_TAX], fabs(ei * axn));
              v[BZ_TZ] = nanmax(v[BZ_TZ], fabs(ei * zn));
              const bool lfin = li > -INF_THRESH, ufin = ui < INF_THRESH;
              double dyp = dy;
              if (!ufin && !lfin) dyp = 0.0;
              else if (!ufin) dyp = fmin(dy, 0.0);
              else if (!lfin) dyp = fmax(dy, 0.0);
              w[row] = dyp;
              v[BZ_NDY] = nanmax(v[BZ_NDY], fabs((E_t ? __ldg(E_t + row) : 1.0) * dyp));
              v[BZ_LHS] += (ufin ? ui * fmax(dyp, 0.0) : 0.0) + (lfin ? li * fmin(dyp, 0.0) : 0.0);
              const double adx = ei * alpha * ad;
              if (ufin) v[BZ_VU] = nanmax(v[BZ_VU], adx);
              if (lfin) v[BZ_VL] = nanmax(v[BZ_VL], -adx);
              v[BZ_SUP] += (ufin ? ui * fmax(yn, 0.0) : 0.0) + (lfin ? li * fmin(yn, 0.0) : 0.0);
            }
          }
        }
        for (int i = tid; i < n; i += NTB) x[i] = fma(alpha, d[i], x[i]);      // _osqp.py:664-668
        if (full_n) {
          cta_reduce<BZ_K, NWB>(v, BZ_MAXMASK, s_red, s_tot);
          s_prim = s_tot[BZ_SPRIM];
          s_nprim = fmax(s_tot[BZ_SAX], s_tot[BZ_SZ]);
          t_prim = s_tot[BZ_TPRIM]; t_ax = s_tot[BZ_TAX]; t_z = s_tot[BZ_TZ];
          ndy = s_tot[BZ_NDY]; lhs = s_tot[BZ_LHS]; vu = s_tot[BZ_VU]; vl = s_tot[BZ_VL]; sup = s_tot[BZ_SUP];
        } else {
          __syncthreads();
        }
      }
      k = kn;
    }

    // ---------------------------------------------------------------- store (reference: _osqp.py:1098-1115)
    __syncthreads();
    const int mode = (status == 3 || status == 4) ? 1 : ((status == 5 || status == 6) ? 2 : 0);
    const bool bad = mode != 0 || status == 9;
    for (int i = tid; i < n; i += NTB) {
      const double Di = __ldg(a.D + i);
      a.xs[on + i] = x[i];
      a.sol_x[on + i] = bad ? NANV : Di * x[i];
      a.cert_n[on + i] = mode == 2 ? Di * alpha * d[i] : NANV;
    }
    for (int i = tid; i < m; i += NTB) {
      const double Ei = __ldg(a.E + i);
      a.ys[om + i] = y[i];
      a.zs[om + i] = z[i];
      a.sol_y[om + i] = bad ? NANV : Ei * y[i] * cinv;
      a.cert_m[om + i] = mode == 1 ? Ei * w[i] : NANV;
    }
    if (tid == 0) {
      BInfo o;
      o.status = status; o.iter = k; o.rho_updates = rho_updates; o.cg_iters = cg_total;
      o.obj = mode == 1 ? QP_INFTY : (mode == 2 ? -QP_INFTY : (status == 9 ? NANV : obj));
      o.dual_obj = mode == 1 ? QP_INFTY : (mode == 2 ? -QP_INFTY : (status == 9 ? NANV : dobj));
      o.prim_res = prim_res; o.dual_res = dual_res; o.rho = rho;
      const double pr = s_prim / (s_nprim + 1e-10), du = s_dual / (s_ndual + 1e-10);
      o.rho_estimate = fmin(fmax(rho * sqrt(pr / (du + 1e-10)), RHO_MIN), RHO_MAX);
      a.info[qp] = o;
      a.rho_b[qp] = rho;
    }
  }
}

}  // namespace b200qp
