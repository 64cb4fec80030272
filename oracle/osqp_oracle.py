"""
CPU ORACLE for the OSQP ADMM solve() path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this file.  The shipped engine (osqp-python_b200/) never does; it
fails loudly when its CUDA library is missing.

What this is
------------
A numpy/scipy restatement of the reference's ADMM algorithm, written from the
reference authors' own pure-Python solver.  Citations "ref:LINE" below are
/root/reference/src/osqppurepy/_osqp.py:LINE.  Two linear-system modes:

* linsys="direct"   -- the reference's formulation: sparse LU of the quasi-definite
                       KKT matrix [[P+sigma I, A'],[A, -diag(1/rho)]]  (ref:281-312).
                       This is the mode that is pinned against the reference.
* linsys="indirect" -- the engine specification: the same ADMM step with nu
                       eliminated, (P+sigma I+A' diag(rho) A) xt = rhs solved by
                       Jacobi-PCG in residual-correction form, tolerance tied to the
                       warm-start residual.  Algebraically identical step
                       (SURVEY.md 8c "reduced-system equivalence"); this mode exists
                       so the CUDA engine can be compared like-for-like
                       (iteration counts, CG counts) and timed as the CPU baseline.

Pinning
-------
direct mode is checked in tests/test_oracle.py against (a) every golden .npz the
reference's tests hold for this path (tests/golden/ref_solutions/*.npz) and
(b) outputs of the reference's osqppurepy run in the build container
(tests/golden/purepy_*.npz, produced by tests/golden/make_golden.py).

Deliberate departures from osqppurepy (bugs listed in SURVEY.md 8c-7, not
reproduced; upstream-C semantics used instead):
  - warm_start scales y by c*Einv (purepy drops c, ref:1506);
  - update of P and A together keeps the cost scale c on P (purepy drops it, ref:1483);
  - check_termination / adaptive_rho_interval are integer cadences (0 = never);
  - "max_iter reached" only downgrades to *_INACCURATE when no status was set
    (purepy re-checks even after SOLVED at the last iteration, ref:1264-1266);
  - m == 0 works with scaling on (purepy crashes on new scipy);
  - status codes use the v1.0 binding order 1..11 (bindings.cpp.in:349-361).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# ---- constants (ref:25-45) --------------------------------------------------
RHO_MIN, RHO_MAX = 1e-6, 1e6
RHO_EQ_FACTOR = 1e3
RHO_TOL = 1e-4
INFTY = 1e30
MIN_SCALING, MAX_SCALING = 1e-4, 1e4
INF_THRESH = INFTY * MIN_SCALING  # "1e26": bound counts as infinite (ref:508-509)

# status values: binding order, bindings.cpp.in:349-361
SOLVED, SOLVED_INACCURATE = 1, 2
PRIMAL_INFEASIBLE, PRIMAL_INFEASIBLE_INACCURATE = 3, 4
DUAL_INFEASIBLE, DUAL_INFEASIBLE_INACCURATE = 5, 6
MAX_ITER_REACHED, TIME_LIMIT_REACHED, NON_CVX, SIGINT, UNSOLVED = 7, 8, 9, 10, 11
STATUS_NAME = {
    SOLVED: "solved", SOLVED_INACCURATE: "solved inaccurate",
    PRIMAL_INFEASIBLE: "primal infeasible",
    PRIMAL_INFEASIBLE_INACCURATE: "primal infeasible inaccurate",
    DUAL_INFEASIBLE: "dual infeasible",
    DUAL_INFEASIBLE_INACCURATE: "dual infeasible inaccurate",
    MAX_ITER_REACHED: "maximum iterations reached",
    TIME_LIMIT_REACHED: "run time limit reached",
    NON_CVX: "problem non convex", SIGINT: "interrupted", UNSOLVED: "unsolved",
}


@dataclass
class Settings:
    """Defaults = the engine's defaults (include/b200qp.h); see ref:145-166 for purepy's."""
    rho: float = 0.1
    rho_is_vec: int = 1
    sigma: float = 1e-6
    alpha: float = 1.6
    scaling: int = 10
    max_iter: int = 4000
    eps_abs: float = 1e-3
    eps_rel: float = 1e-3
    eps_prim_inf: float = 1e-4
    eps_dual_inf: float = 1e-4
    scaled_termination: int = 0
    check_termination: int = 25
    adaptive_rho: int = 1
    adaptive_rho_interval: int = 50
    adaptive_rho_tolerance: float = 5.0
    warm_starting: int = 1
    polishing: int = 0
    delta: float = 1e-6
    polish_refine_iter: int = 3
    time_limit: float = 1e10
    # upstream v1.0 addition (no counterpart in osqppurepy): also require |obj - dual_obj| < eps_abs + eps_rel max(|obj|, |dual_obj|)
    check_dualgap: int = 0
    # indirect mode (engine spec)
    cg_max_iter: int = 100
    cg_tol_fraction: float = 0.1
    cg_precond: int = 1
    # rho on equality rows = rho_eq_factor * rho.  Reference constant: 1e3 (ref:27); the engine's
    # PCG-friendly default is 10 -- pass rho_eq_factor=10 to mirror the engine.
    rho_eq_factor: float = 1e3


# ---- optional multi-threaded SpMV for the engine-spec mode (oracle/spmv_omp.c, built by build.py) --------------
_OMP = None


def _omp_lib():
    """ctypes handle of oracle/_build/liboracle_spmv.so, or False when it was not built."""
    global _OMP
    if _OMP is None:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle_spmv.so")
        try:
            lib = ctypes.CDLL(path)
            lib.oracle_csr_matvec.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 5
            lib.oracle_csr_matvec.restype = None
            lib.oracle_omp_threads.restype = ctypes.c_int
            if hasattr(lib, "oracle_set_threads"):
                lib.oracle_set_threads.argtypes = [ctypes.c_int]
                lib.oracle_set_threads.restype = None
            _OMP = lib
        except OSError:
            _OMP = False
    return _OMP


_OMP_WINS = None      # decided by timing both paths once on the first large product


def host_threads():
    """Threads the oracle's large SpMVs run on: the OpenMP helper's, if it is present AND measured faster than
    scipy's single-threaded csr_matvec on this host (containers with a CPU quota see no gain); else 1."""
    lib = _omp_lib()
    return int(lib.oracle_omp_threads()) if (lib and _OMP_WINS) else 1


def _matvec_omp(lib, M, v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.empty(M.shape[0])
    lib.oracle_csr_matvec(M.shape[0], M.indptr.ctypes.data, M.indices.ctypes.data, M.data.ctypes.data,
                          v.ctypes.data, out.ctypes.data)
    return out


def _matvec(M, v):
    """M @ v for CSR M; large products go through the OpenMP helper when it is available and faster."""
    global _OMP_WINS
    lib = _omp_lib()
    if not lib or M.nnz < 200_000 or M.indices.dtype != np.int32 or M.indptr.dtype != np.int32:
        return M @ v
    if _OMP_WINS is None:
        t = []
        for fn in (lambda: _matvec_omp(lib, M, v), lambda: M @ v):
            fn()
            t0 = time.perf_counter()
            fn(); fn()
            t.append(time.perf_counter() - t0)
        _OMP_WINS = t[0] < 0.8 * t[1]
    return _matvec_omp(lib, M, v) if _OMP_WINS else M @ v


def _limit(v):
    """ref:363-387: norms below MIN_SCALING -> 1, above MAX_SCALING -> MAX_SCALING."""
    v = np.asarray(v, dtype=float)
    out = np.where(v < MIN_SCALING, 1.0, np.minimum(v, MAX_SCALING))
    return out if out.ndim else float(out)


def _rowmax(M):
    """inf-norm of every row of a sparse matrix (0 for empty rows / empty matrix)."""
    if M.shape[0] == 0:
        return np.zeros(0)
    if M.nnz == 0 or M.shape[1] == 0:
        return np.zeros(M.shape[0])
    return np.asarray(abs(M).max(axis=1).todense()).ravel()


def _rowmax_reduceat(M):
    """row-wise max |.| of a CSR matrix with sorted, possibly empty rows."""
    ip = M.indptr
    out = np.zeros(M.shape[0])
    nz = np.diff(ip) > 0
    if nz.any():
        red = np.maximum.reduceat(np.abs(M.data), ip[:-1][nz])
        out[nz] = red
    return out


def full_symmetric(P):
    """The facade sends triu(P) (interface.py:221-222); purepy wants the full matrix."""
    P = sp.csc_matrix(P)
    if sp.tril(P, -1).nnz == 0 and sp.triu(P, 1).nnz > 0:
        P = P + sp.triu(P, 1).T
    return sp.csr_matrix(P)


class _DirectKKT:
    """ref:281-312 -- quasi-definite KKT, sparse LU."""

    def __init__(self, P, A, sigma, rho_inv):
        n, m = P.shape[0], A.shape[0]
        self.n = n
        K = sp.bmat([[P + sigma * sp.eye(n), A.T], [A, -sp.diags(rho_inv)]], format="csc") if m \
            else sp.csc_matrix(P + sigma * sp.eye(n))
        self.lu = spla.splu(K)

    def solve(self, rhs):
        return self.lu.solve(rhs)


class OracleSolver:
    """Same call surface as the path under test: setup / solve / update / warm_start."""

    # ------------------------------------------------------------------ setup
    def __init__(self, P, q, A, l, u, linsys="direct", **kw):
        t0 = time.perf_counter()
        self.s = Settings(**kw)
        self.linsys = linsys
        n = len(q)
        A = sp.csr_matrix(A) if A is not None else sp.csr_matrix((0, n))
        m = A.shape[0]
        self.n, self.m = n, m
        P = full_symmetric(P) if P is not None else sp.csr_matrix((n, n))
        q = np.asarray(q, float).copy()
        l = np.maximum(np.asarray(l, float), -INFTY) if m else np.zeros(0)  # interface.py:237-238
        u = np.minimum(np.asarray(u, float), INFTY) if m else np.zeros(0)
        self.D, self.E, self.c = np.ones(n), np.ones(m), 1.0
        if self.s.scaling:
            P, q, A, l, u = self._ruiz(P, q, A, l, u)
        self.P, self.q, self.A, self.l, self.u = P, q, A, l, u
        self.At = sp.csr_matrix(A.T)
        self.s.rho = float(min(max(self.s.rho, RHO_MIN), RHO_MAX))  # ref:503
        self.ctype = np.zeros(m, dtype=int)
        self.rho_vec = np.zeros(m)
        self._type_constraints()
        self._refresh_linsys()
        self.x, self.z, self.y = np.zeros(n), np.zeros(m), np.zeros(m)
        self.dx, self.dy = np.zeros(n), np.zeros(m)
        self.info = SimpleNamespace(iter=0, status_val=UNSOLVED, status="unsolved", status_polish=0,
                                    obj_val=0.0, dual_obj_val=0.0, duality_gap=0.0, prim_res=0.0, dual_res=0.0, rho_updates=0,
                                    rho_estimate=self.s.rho, setup_time=time.perf_counter() - t0,
                                    solve_time=0.0, polish_time=0.0, cg_iters=0)

    def _ruiz(self, P, q, A, l, u):
        """ref:389-497 (Ruiz equilibration + cost normalisation, `scaling` passes).  Same arithmetic as
        the reference's diag(s) @ M @ diag(s) products, done in place on the CSR value arrays so that the
        n = m = 1e6 baseline problem scales in seconds."""
        n, m = self.n, self.m
        P = sp.csr_matrix(P, dtype=float, copy=True)
        A = sp.csr_matrix(A, dtype=float, copy=True)
        P.sort_indices(); A.sort_indices()
        At = sp.csr_matrix(A.T)                      # column norms of A = row norms of A'
        At.sort_indices()
        prow = np.repeat(np.arange(n), np.diff(P.indptr))
        arow = np.repeat(np.arange(m), np.diff(A.indptr))
        atrow = np.repeat(np.arange(n), np.diff(At.indptr))

        def rowmax(M, rows):
            out = np.zeros(M.shape[0])
            if M.nnz:
                np.maximum.at(out, rows, np.abs(M.data)) if M.nnz < 4096 else out.__setitem__(
                    slice(None), _rowmax_reduceat(M))
            return out

        D, E, c = np.ones(n), np.ones(m), 1.0
        q, l, u = q.copy(), l.copy(), u.copy()
        for _ in range(int(self.s.scaling)):
            # ref:348-361: column norms of the KKT matrix [[P, A'],[A, 0]] (P symmetric: rows = columns)
            colP = rowmax(P, prow)
            colA = rowmax(At, atrow)
            sD = 1.0 / np.sqrt(_limit(np.maximum(colP, colA)))       # ref:419-421
            sE = 1.0 / np.sqrt(_limit(rowmax(A, arow))) if m else np.zeros(0)
            P.data *= sD[prow] * sD[P.indices]                       # ref:432-436
            if m:
                A.data *= sE[arow] * sD[A.indices]
                At.data *= sD[atrow] * sE[At.indices]
            q, l, u = sD * q, sE * l, sE * u
            D, E = sD * D, sE * E                                    # ref:439-440
            # ref:443-448 cost normalisation
            mean_colP = float(rowmax(P, prow).mean()) if P.nnz else 0.0
            qn = _limit(np.max(np.abs(q))) if n else 1.0
            ct = 1.0 / _limit(max(qn, mean_colP))
            P.data *= ct
            q, c = ct * q, ct * c                                    # ref:464-468
        self.D, self.E, self.c = D, E, c
        return P, q, A, l, u

    def _type_constraints(self):
        """ref:499-562: -1 loose, 1 equality, 0 inequality; rho per type."""
        l, u, rho = self.l, self.u, self.s.rho
        loose = (l < -INF_THRESH) & (u > INF_THRESH)
        eq = (u - l) < RHO_TOL
        if not self.s.rho_is_vec:
            self.ctype[:] = 0
            self.rho_vec = np.full(self.m, rho)
        else:
            self.ctype = np.where(eq, 1, np.where(loose, -1, 0))     # eq wins (ref:516-518 order)
            self.rho_vec = np.where(eq, self.s.rho_eq_factor * rho, np.where(loose, RHO_MIN, rho))
        self.rho_inv = 1.0 / self.rho_vec if self.m else np.zeros(0)

    def _refresh_linsys(self):
        if self.linsys == "direct":
            self.kkt = _DirectKKT(self.P, self.A, self.s.sigma, self.rho_inv)
        else:
            # Jacobi preconditioner of the reduced matrix
            diag = self.P.diagonal() + self.s.sigma
            if self.m:
                diag = diag + self.At.multiply(self.At).dot(self.rho_vec)
            self.Minv = 1.0 / diag if self.s.cg_precond else np.ones(self.n)

    def _set_rho(self, rho_new):
        """ref:1579-1597."""
        self.s.rho = float(min(max(rho_new, RHO_MIN), RHO_MAX))
        if self.s.rho_is_vec:
            self.rho_vec = np.where(self.ctype == 1, self.s.rho_eq_factor * self.s.rho,
                                    np.where(self.ctype == -1, RHO_MIN, self.s.rho))
        else:
            self.rho_vec = np.full(self.m, self.s.rho)
        self.rho_inv = 1.0 / self.rho_vec if self.m else np.zeros(0)
        self._refresh_linsys()

    update_rho = _set_rho

    # ---------------------------------------------------------------- one step
    def _kkt_step(self):
        """ref:644-658: returns (xt, zt).  Both modes give zt = A xt up to solve accuracy."""
        s = self.s
        if self.linsys == "direct":
            rhs = np.concatenate([s.sigma * self.x - self.q, self.z - self.rho_inv * self.y])
            sol = self.kkt.solve(rhs)
            xt = sol[: self.n]
            zt = self.z + self.rho_inv * (sol[self.n:] - self.y) if self.m else np.zeros(0)
            return xt, zt
        # ---- engine spec: residual-correction PCG on K = P + sigma I + A' rho A
        # r0 = rhs - K x = -(P x + q) + A'(rho (z - A x) - y);  solve K d = r0, xt = x + d
        ax = _matvec(self.A, self.x)
        r = -(_matvec(self.P, self.x) + self.q)
        if self.m:
            r = r + _matvec(self.At, self.rho_vec * (self.z - ax) - self.y)
        d = np.zeros(self.n)
        r0n = np.max(np.abs(r)) if self.n else 0.0
        tol = s.cg_tol_fraction * r0n
        if r0n > 1e-300:
            zc = self.Minv * r
            p = zc.copy()
            rz = float(r @ zc)
            for _ in range(int(s.cg_max_iter)):
                kp = _matvec(self.P, p) + s.sigma * p
                if self.m:
                    kp = kp + _matvec(self.At, self.rho_vec * _matvec(self.A, p))
                pkp = float(p @ kp)
                if not (pkp > 0.0):
                    self._cg_breakdown = True
                    break
                a = rz / pkp
                d += a * p
                r -= a * kp
                self.info.cg_iters += 1
                zc = self.Minv * r
                rz_new = float(r @ zc)
                if np.max(np.abs(r)) <= tol:
                    break
                p = zc + (rz_new / rz) * p
                rz = rz_new
        xt = self.x + d
        return xt, _matvec(self.A, xt)

    # ---------------------------------------------------------------- residuals
    def _residuals(self, x, z, y):
        """ref:705-794: obj, unscaled-or-scaled residuals and the norms used by the tolerances."""
        s = self.s
        ax, px, aty = self.A @ x, self.P @ x, (self.At @ y if self.m else np.zeros(self.n))
        unscale = bool(s.scaling) and not s.scaled_termination
        Einv = 1.0 / self.E if unscale else np.ones(self.m)
        Dinv = 1.0 / self.D if unscale else np.ones(self.n)
        cinv = 1.0 / self.c if unscale else 1.0
        nrm = lambda v: float(np.max(np.abs(v))) if v.size else 0.0
        r = SimpleNamespace()
        r.obj = (0.5 * float(x @ px) + float(self.q @ x)) / self.c              # ref:705-712
        r.prim = nrm(Einv * (ax - z))                                           # ref:714-726
        r.dual = cinv * nrm(Dinv * (px + self.q + aty))                         # ref:753-764
        r.prim_norm = max(nrm(Einv * ax), nrm(Einv * z))                        # ref:728-751
        r.dual_norm = cinv * max(nrm(Dinv * aty), nrm(Dinv * px), nrm(Dinv * self.q))  # ref:766-794
        # Dual objective and duality gap.  osqppurepy has neither; the binding exposes them (bindings.cpp.in:479,482:
        # dual_obj_val, duality_gap) and upstream v1.0 tests |gap| when check_dualgap is set.  Lagrange dual of the QP:
        #   d(y) = -1/2 x'Px - sup_{l <= z <= u} y'z = -1/2 x'Px - (u'y+ + l'y-)    (finite bounds only; the scaled
        # y, l, u carry c E^-1, E, E, so the products divide by c like the objective).  No reference file pins the
        # numbers; tests check them against this formula, against strong duality at an optimum, and GPU vs oracle.
        fin_u, fin_l = self.u < INF_THRESH, self.l > -INF_THRESH
        sup = float(np.sum(np.where(fin_u, self.u * np.maximum(y, 0), 0.0)) + np.sum(np.where(fin_l, self.l * np.minimum(y, 0), 0.0))) \
            if self.m else 0.0
        r.dual_obj = (-0.5 * float(x @ px) - sup) / self.c
        r.gap = r.obj - r.dual_obj
        # scaled quantities for the rho estimate (ref:880-908)
        r.s_prim, r.s_prim_norm = nrm(ax - z), max(nrm(ax), nrm(z))
        r.s_dual, r.s_dual_norm = nrm(px + self.q + aty), max(nrm(aty), nrm(px), nrm(self.q))
        return r

    def _rho_estimate(self, r):
        pr = r.s_prim / (r.s_prim_norm + 1e-10)
        du = r.s_dual / (r.s_dual_norm + 1e-10)
        return float(min(max(self.s.rho * np.sqrt(pr / (du + 1e-10)), RHO_MIN), RHO_MAX))

    def _primal_infeasible(self, eps):
        """ref:796-820."""
        s = self.s
        unscale = bool(s.scaling) and not s.scaled_termination
        dy = self.dy
        nd = float(np.max(np.abs(self.E * dy if unscale else dy))) if self.m else 0.0
        if nd > eps:
            lhs = float(self.u @ np.maximum(dy, 0) + self.l @ np.minimum(dy, 0))
            if lhs < -eps * nd:
                atdy = self.At @ dy
                if unscale:
                    atdy = atdy / self.D
                return float(np.max(np.abs(atdy))) < eps * nd
        return False

    def _dual_infeasible(self, eps):
        """ref:822-878."""
        s = self.s
        unscale = bool(s.scaling) and not s.scaled_termination
        dx = self.dx
        nd = float(np.max(np.abs(self.D * dx if unscale else dx))) if self.n else 0.0
        cs = self.c if unscale else 1.0
        if nd > eps and float(self.q @ dx) < -cs * eps * nd:
            pdx = self.P @ dx
            if unscale:
                pdx = pdx / self.D
            if float(np.max(np.abs(pdx))) < cs * eps * nd:
                adx = self.A @ dx
                if unscale:
                    adx = adx / self.E
                bad = ((self.u < INF_THRESH) & (adx > eps * nd)) | ((self.l > -INF_THRESH) & (adx < -eps * nd))
                return not bool(bad.any())
        return False

    def _check(self, r, approximate=False):
        """ref:998-1077.  Returns a status value or None."""
        s = self.s
        f = 10.0 if approximate else 1.0
        if r.prim > INFTY or r.dual > INFTY or not np.isfinite(r.prim + r.dual):
            return NON_CVX
        prim_ok = self.m == 0 or r.prim < f * s.eps_abs + f * s.eps_rel * r.prim_norm
        dual_ok = r.dual < f * s.eps_abs + f * s.eps_rel * r.dual_norm
        gap_ok = (not s.check_dualgap) or abs(r.gap) < f * s.eps_abs + f * s.eps_rel * max(abs(r.obj), abs(r.dual_obj))
        if prim_ok and dual_ok and gap_ok:
            return SOLVED_INACCURATE if approximate else SOLVED
        if not prim_ok and self._primal_infeasible(f * s.eps_prim_inf):
            return PRIMAL_INFEASIBLE_INACCURATE if approximate else PRIMAL_INFEASIBLE
        if not dual_ok and self._dual_infeasible(f * s.eps_dual_inf):
            return DUAL_INFEASIBLE_INACCURATE if approximate else DUAL_INFEASIBLE
        return None

    # -------------------------------------------------------------------- solve
    def solve(self, max_wall=None):
        s, info = self.s, self.info
        t0 = time.perf_counter()
        if not s.warm_starting:
            self.x, self.z, self.y = np.zeros(self.n), np.zeros(self.m), np.zeros(self.m)
        info.rho_updates, info.cg_iters, info.status_polish = 0, 0, 0
        self._cg_breakdown = False
        status, r, k = None, None, 0
        for k in range(1, int(s.max_iter) + 1):
            x_prev, z_prev = self.x, self.z
            xt, zt = self._kkt_step()                                            # ref:644-658
            self.x = s.alpha * xt + (1 - s.alpha) * x_prev                       # ref:660-668
            self.dx = self.x - x_prev
            zr = s.alpha * zt + (1 - s.alpha) * z_prev
            self.z = np.minimum(np.maximum(zr + self.rho_inv * self.y, self.l), self.u)  # ref:682-692
            self.dy = self.rho_vec * (zr - self.z)                               # ref:694-703
            self.y = self.y + self.dy
            r = None
            if self._cg_breakdown:
                status = NON_CVX
                r = self._residuals(self.x, self.z, self.y)
                break
            if s.check_termination and k % s.check_termination == 0:
                r = self._residuals(self.x, self.z, self.y)
                status = self._check(r)
                if status is not None:
                    break
            if max_wall is not None and time.perf_counter() - t0 > max_wall:
                status = TIME_LIMIT_REACHED
                break
            if s.adaptive_rho and s.adaptive_rho_interval and k % s.adaptive_rho_interval == 0:
                if r is None:
                    r = self._residuals(self.x, self.z, self.y)
                est = self._rho_estimate(r)                                      # ref:910-930
                info.rho_estimate = est
                if est > s.adaptive_rho_tolerance * s.rho or est < s.rho / s.adaptive_rho_tolerance:
                    self._set_rho(est)
                    info.rho_updates += 1
        if r is None:
            r = self._residuals(self.x, self.z, self.y)
        if status is None:
            status = self._check(r)                                               # ref:1248-1257
        if status is None:
            status = self._check(r, approximate=True) or MAX_ITER_REACHED         # ref:1263-1266
        info.iter, info.status_val, info.status = k, status, STATUS_NAME[status]
        info.obj_val, info.prim_res, info.dual_res = r.obj, r.prim, r.dual
        info.dual_obj_val, info.duality_gap = r.dual_obj, r.gap
        if status in (PRIMAL_INFEASIBLE, PRIMAL_INFEASIBLE_INACCURATE):
            info.obj_val = info.dual_obj_val = INFTY
            info.duality_gap = np.nan
        elif status in (DUAL_INFEASIBLE, DUAL_INFEASIBLE_INACCURATE):
            info.obj_val = info.dual_obj_val = -INFTY
            info.duality_gap = np.nan
        elif status == NON_CVX:
            info.obj_val = info.dual_obj_val = info.duality_gap = np.nan
        info.rho_estimate = self._rho_estimate(r)
        info.solve_time = time.perf_counter() - t0
        if s.polishing and status == SOLVED:
            tp = time.perf_counter()
            self._polish(r)
            info.polish_time = time.perf_counter() - tp
        return self._result()

    def _result(self):
        """ref:1098-1115 + certificates (ref:1065-1066, 1074-1075)."""
        st = self.info.status_val
        nanv = lambda k: np.full(k, np.nan)
        res = SimpleNamespace(info=self.info, x=nanv(self.n), y=nanv(self.m),
                              prim_inf_cert=nanv(self.m), dual_inf_cert=nanv(self.n))
        if st in (PRIMAL_INFEASIBLE, PRIMAL_INFEASIBLE_INACCURATE):
            res.prim_inf_cert = self.E * self.dy
        elif st in (DUAL_INFEASIBLE, DUAL_INFEASIBLE_INACCURATE):
            res.dual_inf_cert = self.D * self.dx
        else:
            res.x = self.D * self.x
            res.y = self.E * self.y / self.c
        return res

    # ------------------------------------------------------------------- polish
    def _polish(self, r_admm):
        """ref:1693-1828: active-set KKT with delta regularisation + iterative refinement."""
        s, n = self.s, self.n
        low = np.where(self.z - self.l < -self.y)[0]                             # ref:1719
        upp = np.where(self.u - self.z < self.y)[0]                              # ref:1720
        act = np.concatenate([low, upp])
        Ar = self.A[act] if self.m else sp.csr_matrix((0, n))
        k = Ar.shape[0]
        Kreg = sp.bmat([[self.P + s.delta * sp.eye(n), Ar.T], [Ar, -s.delta * sp.eye(k)]], format="csc") \
            if k else sp.csc_matrix(self.P + s.delta * sp.eye(n))
        K0 = sp.bmat([[self.P, Ar.T], [Ar, None]], format="csr") if k else sp.csr_matrix(self.P)
        lu = spla.splu(Kreg)
        b = np.concatenate([-self.q, self.l[low], self.u[upp]])
        sol = lu.solve(b)
        for _ in range(int(s.polish_refine_iter)):                               # ref:1699-1707
            sol = sol + lu.solve(b - K0 @ sol)
        px = sol[:n]
        py = np.zeros(self.m)
        py[low] = sol[n:n + len(low)]
        py[upp] = sol[n + len(low):]
        t = (self.A @ px) + py                                                   # ref:676-680
        pz = np.minimum(np.maximum(t, self.l), self.u)
        py = t - pz
        rp = self._residuals(px, pz, py)
        ok = (rp.prim < r_admm.prim and rp.dual < r_admm.dual) or \
             (rp.prim < r_admm.prim and r_admm.dual < 1e-10) or \
             (rp.dual < r_admm.dual and r_admm.prim < 1e-10)                     # ref:1786-1793
        if ok:
            self.x, self.z, self.y = px, pz, py
            self.info.obj_val, self.info.prim_res, self.info.dual_res = rp.obj, rp.prim, rp.dual
            self.info.dual_obj_val, self.info.duality_gap = rp.dual_obj, rp.gap
            self.info.status_polish = 1
        else:
            self.info.status_polish = -1

    # ------------------------------------------------------------------ updates
    def warm_start(self, x=None, y=None):
        """ref:1493-1545 with the c factor on y (inverse of ref:1112)."""
        if x is not None:
            self.x = np.asarray(x, float) / self.D
            self.z = self.A @ self.x
        if y is not None:
            self.y = self.c * np.asarray(y, float) / self.E
        if x is None:
            self.x, self.z = np.zeros(self.n), np.zeros(self.m)
        if y is None:
            self.y = np.zeros(self.m)

    def update(self, q=None, l=None, u=None, Px=None, Px_idx=None, Ax=None, Ax_idx=None,
               P_pattern=None, A_pattern=None):
        """ref:1312-1367, 1431-1491.  Matrix updates need the *unscaled* CSC patterns
        (triu for P) the solver was set up with; values arrive in that CSC order
        (bindings.cpp.in:240-281)."""
        if q is not None:
            self.q = self.c * self.D * np.asarray(q, float)                      # ref:1328
        if l is not None or u is not None:
            ln = self.E * np.maximum(np.asarray(l, float), -INFTY) if l is not None else self.l
            un = self.E * np.minimum(np.asarray(u, float), INFTY) if u is not None else self.u
            if np.any(ln > un):
                raise ValueError("lower bound must not exceed upper bound")     # ref:1348
            self.l, self.u = ln, un
            self._type_constraints()                                             # ref:1364
        if Px is not None:
            Pt = sp.csc_matrix(P_pattern).astype(float).copy()
            Pt.data[np.arange(Pt.nnz) if Px_idx is None else Px_idx] = Px
            self._P_unscaled_triu = Pt
            Pf = full_symmetric(Pt)
            self.P = sp.csr_matrix(self.c * (sp.diags(self.D) @ Pf @ sp.diags(self.D)))   # ref:1443
        if Ax is not None:
            At_ = sp.csc_matrix(A_pattern).astype(float).copy()
            At_.data[np.arange(At_.nnz) if Ax_idx is None else Ax_idx] = Ax
            self.A = sp.csr_matrix(sp.diags(self.E) @ At_ @ sp.diags(self.D))    # ref:1463
            self.At = sp.csr_matrix(self.A.T)
        self._refresh_linsys()
        self.info.status_val, self.info.status = UNSOLVED, "unsolved"


def solve_qp(P, q, A, l, u, linsys="direct", **kw):
    """One-shot helper."""
    return OracleSolver(P, q, A, l, u, linsys=linsys, **kw).solve()


# ---------------------------------------------------------------------------------------------------
# Adjoint derivatives (reference surface: src/osqp/interface.py:519-598, src/bindings.cpp.in:283-319;
# semantics pinned by src/osqp/tests/derivative_test.py: gradients of a loss L(x*, y*) with respect to
# P, q, A, l, u, checked there against finite differences).  The upstream C routine is not vendored; this
# is the textbook active-set adjoint of the KKT system, used as the checker of the engine's implementation.
def adjoint_derivative(P, q, A, l, u, x, y, dx, dy=None, z=None):
    """Given a solution (x, y) and dL/dx (=dx), dL/dy (=dy): returns (dP, dq, dA, dl, du).

    Active rows (polish guess, ref:1719-1720): lower if z - l < -y, upper if u - z < y.  Solve
        [P  Aa'] [rx]   [dx ]
        [Aa  0 ] [ry] = [dya]
    then  dq = -rx,  dP = -(rx x' + x rx')/2 (upper triangle is what the reference compares),
          dA = -(y rx' + ry x'),  dl_i = ry_i on lower-active rows, du_i = ry_i on upper-active rows.
    dP, dA are dense arrays."""
    Pf = full_symmetric(P).toarray() if sp.issparse(P) else np.asarray(P, float)
    Ad = A.toarray() if sp.issparse(A) else np.asarray(A, float)
    n, m = len(q), Ad.shape[0]
    x, y = np.asarray(x, float), np.asarray(y, float)
    dx = np.asarray(dx, float)
    dy = np.zeros(m) if dy is None else np.asarray(dy, float)
    z = Ad @ x if z is None else np.asarray(z, float)
    low = (z - l) < -y
    upp = (u - z) < y
    act = low | upp
    Aa = Ad[act]
    k = int(act.sum())
    K = np.block([[Pf, Aa.T], [Aa, np.zeros((k, k))]])
    sol = np.linalg.lstsq(K, np.concatenate([dx, dy[act]]), rcond=None)[0]
    rx = sol[:n]
    ry = np.zeros(m)
    ry[act] = sol[n:]
    dq = -rx
    dP = -0.5 * (np.outer(rx, x) + np.outer(x, rx))
    dA = -(np.outer(y, rx) + np.outer(ry, x))
    dl = np.where(low, ry, 0.0)
    du = np.where(upp & ~low, ry, 0.0)
    return dP, dq, dA, dl, du
