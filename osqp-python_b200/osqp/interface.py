"""Host-side mirror of the reference facade for the solve() path.

Reference: /src/osqp/interface.py (class OSQP :120, setup :370-397, solve :403-434,
update :330-368, update_settings :280-328, warm_start :399-401, constant :62-89,
SolverStatus/SolverError :97-103, OSQPException :106-117).  Same names, argument meaning and
error behaviour; differences, all deliberate:

  * one algebra: 'cuda' (module `osqp_cuda`).  No multi-backend search, and NO CPU fallback:
    if the extension cannot be imported, or no CUDA device exists at setup(), it raises.
  * codegen is absent (capability bit 0; it emits CPU embedded C).  The adjoint-derivative methods
    (interface.py:519-598) are provided: the adjoint KKT solve runs on device in the same persistent kernel.
"""
import importlib
import os
import warnings
from enum import IntEnum
from types import SimpleNamespace

import numpy as np
import scipy.sparse as spa

_ALGEBRAS = ("cuda",)
_ALGEBRA_MODULES = {"cuda": "osqp_cuda"}
_KNOWN_ELSEWHERE = ("mkl", "builtin")          # reference algebras that do not exist here


def _load(algebra="cuda"):
    if algebra not in _ALGEBRA_MODULES:
        raise RuntimeError(f"Algebra {algebra} not available")
    return importlib.import_module(_ALGEBRA_MODULES[algebra])


def algebra_available(algebra):
    assert algebra in _ALGEBRAS + _KNOWN_ELSEWHERE, f"Unknown algebra {algebra}"
    if algebra not in _ALGEBRA_MODULES:
        return False
    try:
        _load(algebra)
    except ImportError:
        return False
    return True


def algebras_available():
    return [a for a in _ALGEBRAS if algebra_available(a)]


def default_algebra():
    env = os.environ.get("OSQP_ALGEBRA_BACKEND")
    if env is not None:
        return env
    if algebra_available("cuda"):
        return "cuda"
    raise RuntimeError("No algebra backend available! (build osqp_cuda: python __graft_entry__.py)")


def constant(which, algebra="cuda"):
    """Named constant of the extension module (reference interface.py:62-89)."""
    ext = _load(algebra)
    if which in ext.osqp_status_type.__members__:
        warnings.warn("Direct access to osqp status values will be deprecated. "
                      "Please use the SolverStatus enum instead.", PendingDeprecationWarning)
    value = getattr(ext, which, None)
    if isinstance(value, (int, float, str)):
        return value
    if value is not None:
        return value.value
    if which == "OSQP_NAN":
        return np.nan
    raise RuntimeError(f"Unknown constant {which}")


def _int_enum(name, binding_enum):
    ext = _load(default_algebra())
    members = getattr(ext, binding_enum).__members__
    return IntEnum(name, [(k, int(v.value)) for k, v in members.items()])


SolverStatus = _int_enum("SolverStatus", "osqp_status_type")
SolverError = _int_enum("SolverError", "osqp_error_type")


class OSQPException(Exception):
    """Raised when the engine reports an error code; compares equal to that code
    (examples/exception_handling.py:28-33)."""

    def __init__(self, error_code=None):
        super().__init__()
        if error_code:
            self.args = (error_code,)

    def __eq__(self, error_code):
        return len(self.args) > 0 and self.args[0] == error_code

    __hash__ = Exception.__hash__


def _reraise(fn, *args, **kwargs):
    """ValueError(str(code)) from the binding -> OSQPException(code)."""
    try:
        return fn(*args, **kwargs)
    except ValueError as exc:
        code = None
        if exc.args:
            try:
                code = int(exc.args[0])
            except (TypeError, ValueError):
                code = None
        raise OSQPException(code)


_RENAMED = {"polish": "polishing", "warm_start": "warm_starting"}


class OSQP:
    def __init__(self, *args, **kwargs):
        self.m = None
        self.n = None
        self.algebra = kwargs.pop("algebra") if "algebra" in kwargs else default_algebra()
        if self.algebra not in _ALGEBRA_MODULES or not algebra_available(self.algebra):
            raise RuntimeError(f"Algebra {self.algebra} not available")
        self.ext = _load(self.algebra)
        self._dtype = np.float32 if self.ext.OSQP_USE_FLOAT == 1 else np.float64
        self._itype = np.int64 if self.ext.OSQP_USE_LONG == 1 else np.int32
        self._solver = None
        self.settings = None
        self._derivative_cache = {}

    raises_error = staticmethod(_reraise)

    def __str__(self):
        if self._solver is None:
            return f"Uninitialized OSQP with algebra={self.algebra}"
        return f"OSQP with algebra={self.algebra} ({self.solver_type})"

    # ------------------------------------------------------------------ input normalisation
    def _infer_mnpqalu(self, P=None, q=None, A=None, l=None, u=None):
        """Sizes, defaults and canonical formats (reference interface.py:165-240): P -> upper
        triangular CSC, A -> CSC, sorted indices, bounds clipped to +-OSQP_INFTY."""
        if P is not None:
            n = P.shape[0]
        elif q is not None:
            n = len(q)
        elif A is not None:
            n = A.shape[1]
        else:
            raise ValueError("The problem does not have any variables")
        m = 0 if A is None else A.shape[0]

        if A is None:
            assert l is None and u is None, "If A is unspecified, leave l/u unspecified too."
        else:
            assert l is not None or u is not None, "If A is specified, specify at least one of l/u."
            l = np.full(m, -np.inf) if l is None else l
            u = np.full(m, np.inf) if u is None else u

        def empty(rows):
            return spa.csc_matrix((np.zeros(0, self._dtype), np.zeros(0, self._itype), np.zeros(n + 1, self._itype)),
                                  shape=(rows, n))

        if P is None:
            P = empty(n)
        if q is None:
            q = np.zeros(n)
        if A is None:
            A, l, u = empty(0), np.zeros(0), np.zeros(0)
        q, l, u = np.asarray(q), np.asarray(l), np.asarray(u)
        assert len(q) == n, "Incorrect dimension of q"
        assert len(l) == m, "Incorrect dimension of l"
        assert len(u) == m, "Incorrect dimension of u"

        for name, M in (("P", P), ("A", A)):
            if not spa.issparse(M) and isinstance(M, np.ndarray) and M.ndim == 2:
                raise TypeError(f"{name} is required to be a sparse matrix")
        if spa.tril(P, -1).data.size > 0:
            P = spa.triu(P, format="csc")
        if not spa.isspmatrix_csc(P):
            warnings.warn("Converting sparse P to a CSC matrix. This may take a while...")
            P = P.tocsc()
        if not spa.isspmatrix_csc(A):
            warnings.warn("Converting sparse A to a CSC matrix. This may take a while...")
            A = A.tocsc()
        for M in (P, A):
            if not M.has_sorted_indices:
                M.sort_indices()
        infty = self.constant("OSQP_INFTY")
        return m, n, P, q, A, np.maximum(l, -infty), np.minimum(u, infty)

    # ------------------------------------------------------------------ introspection
    @property
    def capabilities(self):
        return int(self.ext.osqp_capabilities())

    def has_capability(self, capability: str):
        try:
            cap = int(self.ext.osqp_capabilities_type.__members__[capability])
        except KeyError:
            raise RuntimeError(f"Unrecognized capability {capability}")
        return (self.capabilities & cap) != 0

    @property
    def solver_type(self):
        direct = self.ext.osqp_linsys_solver_type.OSQP_DIRECT_SOLVER
        return "direct" if self.settings.linsys_solver == direct else "indirect"

    @property
    def cg_preconditioner(self):
        return "diagonal" if self.settings.cg_precond == self.ext.OSQP_DIAGONAL_PRECONDITIONER else None

    def constant(self, which):
        return constant(which, algebra=self.algebra)

    # ------------------------------------------------------------------ settings
    def update_settings(self, **kwargs):
        assert self.settings is not None
        for old, new in _RENAMED.items():
            if old in kwargs:
                warnings.warn(f'"{old}" is deprecated. Please use "{new}" instead.', DeprecationWarning)
                kwargs[new] = kwargs.pop(old)

        changed = False
        if "rho" in kwargs and self._solver is not None:
            self._solver.update_rho(kwargs.pop("rho"))
        if "solver_type" in kwargs:
            value = kwargs.pop("solver_type")
            assert value in ("direct", "indirect")
            kinds = self.ext.osqp_linsys_solver_type
            self.settings.linsys_solver = kinds.OSQP_DIRECT_SOLVER if value == "direct" else kinds.OSQP_INDIRECT_SOLVER
            changed = True
        if "cg_preconditioner" in kwargs:
            value = kwargs.pop("cg_preconditioner")
            assert value in (None, "diagonal")
            self.settings.cg_precond = (self.ext.OSQP_DIAGONAL_PRECONDITIONER if value == "diagonal"
                                        else self.ext.OSQP_NO_PRECONDITIONER)
            changed = True
        for field in vars(self.ext.OSQPSettings):
            if not field.startswith("_") and field in kwargs:
                setattr(self.settings, field, kwargs.pop(field))
                changed = True
        if kwargs:
            raise ValueError(f"Unrecognized settings {list(kwargs.keys())}")
        if changed and self._solver is not None:
            _reraise(self._solver.update_settings, self.settings)

    # ------------------------------------------------------------------ the path
    def setup(self, P, q, A, l, u, **settings):
        m, n, P, q, A, l, u = self._infer_mnpqalu(P=P, q=q, A=A, l=l, u=u)
        self.m, self.n = m, n
        Pc = self.ext.CSC(P.astype(self._dtype))
        Ac = self.ext.CSC(A.astype(self._dtype))
        q, l, u = (np.ascontiguousarray(v, dtype=self._dtype) for v in (q, l, u))

        self.settings = self.ext.OSQPSettings()
        self.ext.osqp_set_default_settings(self.settings)
        self._solver = None
        self.update_settings(**settings)
        self._solver = _reraise(self.ext.OSQPSolver, Pc, q, Ac, l, u, m, n, self.settings)
        if "rho" in settings:
            self._solver.update_rho(settings["rho"])
        # problem data as the solver holds it (P upper triangular), for the derivative methods (reference :372)
        self._derivative_cache = {"P": P.astype(self._dtype), "q": q, "A": A.astype(self._dtype), "l": l, "u": u}

    def update(self, **kwargs):
        """q / l / u and/or Px, Px_idx, Ax, Ax_idx (reference interface.py:330-368)."""
        q, l, u = kwargs.get("q"), kwargs.get("l"), kwargs.get("u")
        infty = self.constant("OSQP_INFTY")
        if l is not None:
            l = np.maximum(l, -infty)
        if u is not None:
            u = np.minimum(u, infty)
        if q is not None or l is not None or u is not None:
            self._solver.update_data_vec(q=q, l=l, u=u)
        if any(k in kwargs for k in ("Px", "Px_idx", "Ax", "Ax_idx")):
            self._solver.update_data_mat(P_x=kwargs.get("Px"), P_i=kwargs.get("Px_idx"),
                                         A_x=kwargs.get("Ax"), A_i=kwargs.get("Ax_idx"))
        # keep the cached problem data current and forget the last solution (reference :348-368)
        for name, val in (("q", q), ("l", l), ("u", u)):
            if val is not None and name in self._derivative_cache:
                self._derivative_cache[name] = np.asarray(val, dtype=self._dtype)
        for name, key in (("P", "Px"), ("A", "Ax")):
            if kwargs.get(key) is not None and name in self._derivative_cache:
                idx = kwargs.get(key + "_idx")
                if idx is None:
                    self._derivative_cache[name].data[:] = kwargs[key]
                else:
                    self._derivative_cache[name].data[idx] = kwargs[key]
        self._derivative_cache.pop("results", None)

    def warm_start(self, x=None, y=None):
        return self._solver.warm_start(x, y)

    def solve(self, raise_error=None):
        if raise_error is None:
            warnings.warn("The default value of raise_error will change to True in the future.",
                          PendingDeprecationWarning)
            raise_error = False
        rc = self._solver.solve()
        if rc:
            raise OSQPException(rc)          # engine failure (e.g. lost device); never a silent fallback
        info = self._solver.info
        if info.status_val == SolverStatus.OSQP_NON_CVX:
            info.obj_val = np.nan
        if info.status_val != SolverStatus.OSQP_SOLVED and raise_error:
            raise OSQPException(info.status_val)
        fields = {k: getattr(info, k) for k in vars(type(info)) if not k.startswith("_")}
        sol = self._solver.solution
        results = SimpleNamespace(x=sol.x, y=sol.y, prim_inf_cert=sol.prim_inf_cert, dual_inf_cert=sol.dual_inf_cert,
                                  info=SimpleNamespace(**fields))
        self._derivative_cache["results"] = results
        return results

    # engine counters of the last solve (not in the reference; used by bench.py)
    def stats(self):
        return self._solver.get_stats()

    # ------------------------------------------------------------------ out of scope
    def codegen(self, *a, **k):
        raise NotImplementedError("codegen is out of scope of the B200 engine (OSQP_CAPABILITY_CODEGEN = 0)")

    # ------------------------------------------------------------------ adjoint derivatives (reference :519-598)
    def _solved_results(self):
        assert self.has_capability("OSQP_CAPABILITY_DERIVATIVES"), "This OSQP object does not support derivatives"
        try:
            results = self._derivative_cache["results"]
        except KeyError:
            raise ValueError("Problem has not been solved. You cannot take derivatives. Please call the solve function.")
        if results.info.status_val != SolverStatus.OSQP_SOLVED:
            raise ValueError("Problem has not been solved to optimality. You cannot take derivatives")
        return results

    def adjoint_derivative_compute(self, dx=None, dy=None):
        """Adjoint derivative after solve(): dx = dL/dx, dy = dL/dy (reference interface.py:519-540)."""
        self._solved_results()
        dx = np.zeros(self.n) if dx is None else np.ascontiguousarray(dx, dtype=self._dtype)
        dy = np.zeros(self.m) if dy is None else np.ascontiguousarray(dy, dtype=self._dtype)
        rc = self._solver.adjoint_derivative_compute(dx, dy)
        if rc:
            raise OSQPException(rc)

    def adjoint_derivative_get_mat(self, as_dense=True, dP_as_triu=True):
        """dP / dA after adjoint_derivative_compute (reference interface.py:542-573)."""
        self._solved_results()
        P, A = self._derivative_cache["P"], self._derivative_cache["A"]
        dPm, dAm = P.copy(), A.copy()
        rc = self._solver.adjoint_derivative_get_mat(self.ext.CSC(dPm), self.ext.CSC(dAm))
        if rc:
            raise OSQPException(rc)
        if not dP_as_triu:
            dPm = (dPm + spa.triu(dPm, 1).T).tocsc()
        if as_dense:
            return np.array(dPm.todense()), np.array(dAm.todense())
        return dPm, dAm

    def adjoint_derivative_get_vec(self):
        """dq / dl / du after adjoint_derivative_compute (reference interface.py:575-598)."""
        self._solved_results()
        dq = np.empty(self.n, dtype=self._dtype)
        dl = np.zeros(self.m, dtype=self._dtype)
        du = np.zeros(self.m, dtype=self._dtype)
        rc = self._solver.adjoint_derivative_get_vec(dq, dl, du)
        if rc:
            raise OSQPException(rc)
        return dq, dl, du
