"""Batched front end: many QPs that share (P, A) and differ in q, l, u, solved in one kernel launch.

Reference: the only batched caller in the reference is src/osqp/nn/torch.py:128-164, 200-217 -- it keeps one
solver per batch element, re-uses the setup through update() (:136-140) and loops solve() over a thread
pool.  OSQPBatch is that loop as one object: element b is the QP of setup(P, q[0], A, l[0], u[0]) followed by
update(q=q[b], l=l[b], u=u[b]) and solve().  Same method names and argument meaning as osqp.OSQP
(src/osqp/interface.py:330-434); every array gains a leading batch axis.  BASELINE.json configs[4].
"""
from types import SimpleNamespace

import numpy as np

from .interface import OSQP, OSQPException, SolverStatus, _reraise

_STATUS_TEXT = {
    1: "solved", 2: "solved inaccurate", 3: "primal infeasible", 4: "primal infeasible inaccurate",
    5: "dual infeasible", 6: "dual infeasible inaccurate", 7: "maximum iterations reached",
    8: "run time limit reached", 9: "problem non convex", 10: "interrupted", 11: "unsolved",
}


class OSQPBatch(OSQP):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.batch = None

    def _batch_array(self, v, width, name):
        """(batch, width) float64 C-contiguous; a 1-D array of length `width` is broadcast over the batch."""
        v = np.asarray(v, dtype=self._dtype)
        if v.ndim == 1:
            assert len(v) == width, f"Incorrect dimension of {name}"
            v = np.broadcast_to(v, (self.batch, width))
        assert v.shape == (self.batch, width), f"Incorrect shape of {name}: {v.shape}, expected {(self.batch, width)}"
        return np.ascontiguousarray(v)

    def setup(self, P, q, A, l, u, batch=None, Px=None, Ax=None, **settings):
        """P, A: the shared sparsity pattern (and the values of element 0, which fix the scaling); q (batch, n) or (n,),
        l / u (batch, m) or (m,).  Px / Ax: optional per-element values, see update()."""
        q = np.asarray(q, dtype=np.float64)
        if batch is None:
            assert q.ndim == 2 or np.ndim(l) == 2 or np.ndim(u) == 2, "give batch= or 2-D q / l / u"
            batch = next(np.shape(v)[0] for v in (q, l, u) if np.ndim(v) == 2)
        self.batch = int(batch)
        first = lambda v: None if v is None else (np.asarray(v)[0] if np.ndim(v) == 2 else np.asarray(v))
        m, n, Pn, _, An, _, _ = self._infer_mnpqalu(P=P, q=first(q), A=A, l=first(l), u=first(u))
        self.m, self.n = m, n
        if m:
            l = np.full(m, -np.inf) if l is None else l
            u = np.full(m, np.inf) if u is None else u
            lb = self._batch_array(l, m, "l")                                  # clipped to +-OSQP_INFTY on device
            ub = self._batch_array(u, m, "u")
        else:
            lb = ub = np.zeros((self.batch, 0))
        qb = self._batch_array(q, n, "q")
        Pc = self.ext.CSC(Pn.astype(self._dtype))
        Ac = self.ext.CSC(An.astype(self._dtype))
        self.settings = self.ext.OSQPSettings()
        self.ext.osqp_set_default_settings(self.settings)
        self._solver = None
        self.update_settings(**settings)
        self._solver = _reraise(self.ext.OSQPBatchSolver, Pc, qb, Ac, lb, ub, self.batch, m, n, self.settings)
        if "rho" in settings:
            self._solver.update_rho(settings["rho"])
        if Px is not None or Ax is not None:
            self.update(Px=Px, Ax=Ax)

    def update(self, **kwargs):
        """q / l / u and / or Px / Ax with a leading batch axis (reference interface.py:330-368).  Px is
        (batch, nnz(triu P)), Ax is (batch, nnz(A)), values in the CSC order of the matrices given at setup -- the
        per-element `P_val` / `A_val` of the reference's batched caller (nn/torch.py:136-140).  Index subsets
        (Px_idx / Ax_idx) are not supported in a batch."""
        if "Px_idx" in kwargs or "Ax_idx" in kwargs:
            raise NotImplementedError("OSQPBatch.update takes full value arrays (no Px_idx / Ax_idx)")
        Px, Ax = kwargs.get("Px"), kwargs.get("Ax")
        if Px is not None or Ax is not None:
            Px = None if Px is None else np.ascontiguousarray(np.broadcast_to(np.asarray(Px, dtype=self._dtype), (self.batch, np.shape(Px)[-1])))
            Ax = None if Ax is None else np.ascontiguousarray(np.broadcast_to(np.asarray(Ax, dtype=self._dtype), (self.batch, np.shape(Ax)[-1])))
            rc = _reraise(self._solver.update_data_mat, Px, Ax)
            if rc:
                raise OSQPException(rc)
        q, l, u = kwargs.get("q"), kwargs.get("l"), kwargs.get("u")
        q = None if q is None else self._batch_array(q, self.n, "q")
        l = None if l is None else self._batch_array(l, self.m, "l")      # clipped to +-OSQP_INFTY on device
        u = None if u is None else self._batch_array(u, self.m, "u")
        if q is not None or l is not None or u is not None:
            rc = self._solver.update_data_vec(q=q, l=l, u=u)
            if rc:
                raise OSQPException(rc)

    def warm_start(self, x=None, y=None):
        x = None if x is None else self._batch_array(x, self.n, "x")
        y = None if y is None else self._batch_array(y, self.m, "y")
        return self._solver.warm_start(x, y)

    def solve(self, raise_error=False):
        rc = self._solver.solve()
        if rc:
            raise OSQPException(rc)
        info = self._solver.get_info()
        sv = info["status_val"]
        # certificates exist only for infeasible elements: fetch them only then (None otherwise)
        infeasible = bool(np.any((sv >= SolverStatus.OSQP_PRIMAL_INFEASIBLE) & (sv <= SolverStatus.OSQP_DUAL_INFEASIBLE_INACCURATE)))
        sol = self._solver.get_solution(certificates=infeasible)
        if raise_error and np.any(sv != SolverStatus.OSQP_SOLVED):
            raise OSQPException(int(sv[sv != SolverStatus.OSQP_SOLVED][0]))
        info["status"] = [_STATUS_TEXT.get(int(s), "unsolved") for s in sv]
        return SimpleNamespace(x=sol["x"], y=sol["y"], prim_inf_cert=sol.get("prim_inf_cert"),
                               dual_inf_cert=sol.get("dual_inf_cert"), info=SimpleNamespace(**info))

    def stats(self):
        return self._solver.get_stats()

    # one adjoint solve per element is a per-solver operation: use osqp.OSQP for derivatives
    def adjoint_derivative_compute(self, *a, **k):
        raise NotImplementedError("adjoint derivatives are per solver: use osqp.OSQP (one object per QP)")

    adjoint_derivative_get_mat = adjoint_derivative_compute
    adjoint_derivative_get_vec = adjoint_derivative_compute
