// aux_kernels.cuh -- setup / update / solution kernels around the persistent ADMM kernel.
// None of these is on the per-iteration path; they run once per setup(), update() or solve().
#pragma once
#include "device_common.cuh"

namespace b200qp {

// ---- Ruiz equilibration (reference: _osqp.py:389-497) -------------------------------------------
// inf-norm of every CSR row (8 lanes per row)
__global__ void k_row_absmax(CsrMat M, double* __restrict__ out) {
  const int lane = threadIdx.x & 7;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  double v = 0.0;
  if (row < M.rows) {
    const int a = M.ptr[row], b = M.ptr[row + 1];
    for (int k = a + lane; k < b; k += 8) v = fmax(v, fabs(M.val[k]));
  }
#pragma unroll
  for (int off = 4; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
  if (row < M.rows && lane == 0) out[row] = v;
}

__device__ __forceinline__ double limit_scaling(double v) {   // _osqp.py:363-387
  return v < MIN_SCALING ? 1.0 : (v > MAX_SCALING ? MAX_SCALING : v);
}

// s = 1/sqrt(limit(max(a, b)));  acc *= s;  (b may be null)
__global__ void k_ruiz_factor(int len, const double* na, const double* nb, double* s, double* acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len) {
    double v = na ? na[i] : 0.0;
    if (nb) v = fmax(v, nb[i]);
    const double f = 1.0 / sqrt(limit_scaling(v));
    s[i] = f;
    acc[i] *= f;
  }
}

// val[k] *= scalar * rs[row] * cs[col]   (rs / cs may be null = 1)
__global__ void k_scale_csr(int rows, const int* __restrict__ ptr, const int* __restrict__ idx,
                            double* __restrict__ val, const double* __restrict__ rs,
                            const double* __restrict__ cs, double scalar) {
  const int lane = threadIdx.x & 7;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  if (row < rows) {
    const double r = scalar * (rs ? rs[row] : 1.0);
    const int a = ptr[row], b = ptr[row + 1];
    for (int k = a + lane; k < b; k += 8) val[k] *= r * (cs ? cs[idx[k]] : 1.0);
  }
}

__global__ void k_vec_mul(int len, double* __restrict__ v, const double* __restrict__ s, double scalar) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len) v[i] *= scalar * (s ? s[i] : 1.0);
}
__global__ void k_vec_set_mul(int len, double* __restrict__ out, const double* __restrict__ in,
                              const double* __restrict__ s, double scalar) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len) out[i] = scalar * in[i] * (s ? s[i] : 1.0);
}
__global__ void k_vec_recip(int len, double* __restrict__ out, const double* __restrict__ in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len) out[i] = 1.0 / in[i];
}
__global__ void k_vec_fill(int len, double* __restrict__ out, double v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len) out[i] = v;
}

// two-stage reduction: out[0] = max|v.*s|, out[1] = sum v   (fixed order -> deterministic)
constexpr int RED_BLOCKS = 256;
__global__ void k_absmax_sum_partial(int len, const double* __restrict__ v, const double* __restrict__ s,
                                     double* __restrict__ part /* [2*RED_BLOCKS] */) {
  __shared__ double sm[32], ss[32];
  double mx = 0.0, sum = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x) {
    const double x = v[i];
    mx = nanmax(mx, fabs(x * (s ? s[i] : 1.0)));
    sum += x;
  }
  mx = warp_combine(mx, true);
  sum = warp_combine(sum, false);
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = mx; ss[threadIdx.x >> 5] = sum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    for (int w = 1; w < nw; ++w) { mx = nanmax(mx, sm[w]); sum += ss[w]; }
    part[blockIdx.x] = mx; part[RED_BLOCKS + blockIdx.x] = sum;
  }
}
__global__ void k_absmax_sum_final(int nblocks, const double* __restrict__ part, double* __restrict__ out) {
  if (threadIdx.x == 0) {
    double mx = part[0], sum = part[RED_BLOCKS];
    for (int b = 1; b < nblocks; ++b) { mx = nanmax(mx, part[b]); sum += part[RED_BLOCKS + b]; }
    out[0] = mx; out[1] = sum;
  }
}

// read a buffer (L2 flush by reads / bringing a vector back into L2); the result is never true, `sink` is never written
__global__ void k_touch(const double* __restrict__ p, long long len, double* sink) {
  double a = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x) a += __ldcg(p + i);
  if (a == -1.2345e300) sink[0] = a;
}

// values of a scaled CSR matrix -> their slots in the compacted-slice tile blob (csl_host.h)
__global__ void k_csl_pack(int cnt, const int* __restrict__ src, const unsigned* __restrict__ dst,
                           const double* __restrict__ val, double* __restrict__ blob) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < cnt) blob[dst[e]] = val[src[e]];
}

// ---- rho vector / constraint typing (reference: _osqp.py:499-562) -------------------------------
__global__ void k_type_and_rho(int m, const double* __restrict__ l, const double* __restrict__ u, double rho,
                               int rho_is_vec, double eq_factor, int* __restrict__ ctype, double* __restrict__ rho_v,
                               double* __restrict__ rho_inv, int retype) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) {
    int ct = ctype[i];
    if (retype) {
      const double li = l[i], ui = u[i];
      ct = (ui - li < RHO_TOL) ? 1 : ((li < -INF_THRESH && ui > INF_THRESH) ? -1 : 0);
      ctype[i] = ct;
    }
    const double rv = rho_is_vec ? (ct == 1 ? eq_factor * rho : (ct == -1 ? RHO_MIN : rho)) : rho;
    rho_v[i] = rv;
    rho_inv[i] = 1.0 / rv;
  }
}

// Jacobi preconditioner of K = P + sigma I + A' diag(rho) A over the rows of A' 
__global__ void k_jacobi(CsrMat At, const double* __restrict__ pdiag, const double* __restrict__ rho_v,
                         double sigma, int enabled, double* __restrict__ Minv) {
  const int lane = threadIdx.x & 7;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  double v = 0.0;
  if (row < At.rows && enabled) {
    const int a = At.ptr[row], b = At.ptr[row + 1];
    for (int k = a + lane; k < b; k += 8) { const double t = At.val[k]; v = fma(t * t, rho_v[At.idx[k]], v); }
  }
#pragma unroll
  for (int off = 4; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if (row < At.rows && lane == 0) Minv[row] = enabled ? 1.0 / (pdiag[row] + sigma + v) : 1.0;
}

// generic y = M x (8 lanes per row); used by warm start and polish bookkeeping
__global__ void k_spmv_simple(CsrMat M, const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 7;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  double v = 0.0;
  if (row < M.rows) {
    const int a = M.ptr[row], b = M.ptr[row + 1];
    for (int k = a + lane; k < b; k += 8) v = fma(M.val[k], x[M.idx[k]], v);
  }
#pragma unroll
  for (int off = 4; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if (row < M.rows && lane == 0) y[row] = v;
}

// ---- matrix value updates (reference: _osqp.py:1431-1491; bindings.cpp.in:240-281) ---------------
// New values arrive UNSCALED in the CSC order of A; position k of that order is position k of the
// CSR(A') arrays (same memory layout) and position mapA[k] of CSR(A).
__global__ void k_update_A(int cnt, const double* __restrict__ vals, const int* __restrict__ sel,
                           const int* __restrict__ colptr, const int* __restrict__ rowidx, int ncols,
                           const int* __restrict__ mapA, const double* __restrict__ D,
                           const double* __restrict__ E, double* __restrict__ Atx, double* __restrict__ Ax) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt) return;
  const int k = sel ? sel[t] : t;
  int lo = 0, hi = ncols;               // column j with colptr[j] <= k < colptr[j+1]
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (colptr[mid] <= k) lo = mid; else hi = mid; }
  const double v = E[rowidx[k]] * vals[t] * D[lo];
  Atx[k] = v;
  Ax[mapA[k]] = v;
}
// P arrives as the upper triangle in CSC order; both mirrored positions of the full CSR are written.
__global__ void k_update_P(int cnt, const double* __restrict__ vals, const int* __restrict__ sel,
                           const int* __restrict__ colptr, const int* __restrict__ rowidx, int ncols,
                           const int* __restrict__ posU, const int* __restrict__ posL,
                           const double* __restrict__ D, double c, double* __restrict__ Px,
                           double* __restrict__ pdiag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt) return;
  const int k = sel ? sel[t] : t;
  int lo = 0, hi = ncols;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (colptr[mid] <= k) lo = mid; else hi = mid; }
  const int i = rowidx[k];
  const double v = c * D[i] * vals[t] * D[lo];
  Px[posU[k]] = v;
  if (posL[k] >= 0) Px[posL[k]] = v; else pdiag[i] = v;
}

// ---- warm start (reference: _osqp.py:1493-1545; y gets the cost scale, inverse of :1112) ----------
__global__ void k_warm_x(int n, const double* __restrict__ xin, const double* __restrict__ Dinv,
                         double* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = xin ? xin[i] * Dinv[i] : 0.0;
}
__global__ void k_warm_y(int m, const double* __restrict__ yin, const double* __restrict__ Einv, double c,
                         double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) y[i] = yin ? c * yin[i] * Einv[i] : 0.0;
}

// ---- solution (reference: _osqp.py:1098-1115, certificates :1065-1066, :1074-1075) ---------------
// mode 0: x = D x, y = E y / c ; 1: prim_inf_cert = E * dy_projected ; 2: dual_inf_cert = D * alpha d
__global__ void k_store_n(int n, int mode, const double* __restrict__ x, const double* __restrict__ d,
                          const double* __restrict__ D, double alpha, double* __restrict__ out_x,
                          double* __restrict__ out_dcert) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double nanv = __longlong_as_double(0x7ff8000000000000LL);
  out_x[i] = mode == 0 ? D[i] * x[i] : nanv;
  out_dcert[i] = mode == 2 ? D[i] * alpha * d[i] : nanv;
}
__global__ void k_store_m(int m, int mode, const double* __restrict__ y, const double* __restrict__ dyp,
                          const double* __restrict__ E, double cinv, double* __restrict__ out_y,
                          double* __restrict__ out_pcert) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double nanv = __longlong_as_double(0x7ff8000000000000LL);
  out_y[i] = mode == 0 ? E[i] * y[i] * cinv : nanv;
  out_pcert[i] = mode == 1 ? E[i] * dyp[i] : nanv;
}

// ---- polish support (reference: _osqp.py:1710-1828) ------------------------------------------------
// Active-set guess (ref:1719-1720) and the bounds / rho / duals of the equality-constrained problem the
// persistent kernel is re-run on: active rows become equalities at the active bound, the others are
// released (infinite bounds, rho = RHO_MIN, y = 0).
__global__ void k_polish_prepare(int m, const double* __restrict__ l, const double* __restrict__ u,
                                 const double* __restrict__ z, const double* __restrict__ ax, double rho_act,
                                 double* __restrict__ y, double* __restrict__ zp, double* __restrict__ lp,
                                 double* __restrict__ up, int* __restrict__ ctp, double* __restrict__ rho_v,
                                 double* __restrict__ rho_inv, int* __restrict__ n_active) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double zi = z[i], yi = y[i], li = l[i], ui = u[i];
  const bool low = (zi - li) < -yi, upp = (ui - zi) < yi;
  if (low || upp) {
    const double b = low ? li : ui;
    lp[i] = b; up[i] = b; zp[i] = b; ctp[i] = 1;
    rho_v[i] = rho_act; rho_inv[i] = 1.0 / rho_act;
    atomicAdd(n_active, 1);
  } else {
    lp[i] = -QP_INFTY; up[i] = QP_INFTY; zp[i] = ax[i]; ctp[i] = -1;
    rho_v[i] = RHO_MIN; rho_inv[i] = 1.0 / RHO_MIN;
    y[i] = 0.0;
  }
}

// (z, y) <- normal-cone projection of (A x + y) on the ORIGINAL box (ref:676-680); partial: |Einv (Ax - z)|inf
__global__ void k_polish_normalcone(int m, const double* __restrict__ ax, const double* __restrict__ l,
                                    const double* __restrict__ u, const double* __restrict__ Einv,
                                    double* __restrict__ y, double* __restrict__ z, double* __restrict__ res) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double t = ax[i] + y[i];
  const double zn = fmin(fmax(t, l[i]), u[i]);
  z[i] = zn;
  y[i] = t - zn;
  res[i] = (Einv ? Einv[i] : 1.0) * (ax[i] - zn);
}
// dres = Dinv .* (px + q + aty) ; xq = x .* (0.5 px + q)
__global__ void k_polish_dual(int n, const double* __restrict__ px, const double* __restrict__ q,
                              const double* __restrict__ aty, const double* __restrict__ Dinv,
                              const double* __restrict__ x, double* __restrict__ dres, double* __restrict__ xq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dres[i] = (Dinv ? Dinv[i] : 1.0) * (px[i] + q[i] + aty[i]);
  xq[i] = x[i] * (0.5 * px[i] + q[i]);
}

// ---- adjoint derivatives (reference surface: interface.py:519-598) ---------------------------------------------
// Auxiliary equality-constrained QP of the adjoint KKT system: active rows (same guess as polish, ref:1719-1720)
// become equalities a_i'r = dy_i (scaled: E_i dy_i), the others are released; act = 1 lower-active, 2 upper-active.
__global__ void k_adjoint_prepare(int m, const double* __restrict__ l, const double* __restrict__ u,
                                  const double* __restrict__ z, const double* __restrict__ y,
                                  const double* __restrict__ dy, const double* __restrict__ E, double rho_act,
                                  double* __restrict__ lp, double* __restrict__ up, int* __restrict__ ctp,
                                  double* __restrict__ rho_v, double* __restrict__ rho_inv, int* __restrict__ act) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double zi = z[i], yi = y[i];
  const bool low = (zi - l[i]) < -yi, upp = (u[i] - zi) < yi;
  act[i] = low ? 1 : (upp ? 2 : 0);
  if (low || upp) {
    const double b = dy ? E[i] * dy[i] : 0.0;
    lp[i] = b; up[i] = b; ctp[i] = 1;
    rho_v[i] = rho_act; rho_inv[i] = 1.0 / rho_act;
  } else {
    lp[i] = -QP_INFTY; up[i] = QP_INFTY; ctp[i] = -1;
    rho_v[i] = RHO_MIN; rho_inv[i] = 1.0 / RHO_MIN;
  }
}

}  // namespace b200qp
