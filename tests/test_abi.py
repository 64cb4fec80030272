"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/b200qp.h declares, the pybind11 module exposes the surface the reference facade consumes
(SURVEY.md 8b), and -- without a GPU -- the product path fails loudly instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200qp.h")
LIB = os.path.join(ROOT, "osqp-python_b200", "lib", "libb200qp.so")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(b200qp_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build first: python __graft_entry__.py"
    lib = ctypes.CDLL(LIB)
    names = _declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/b200qp.h but not exported"


def test_default_settings_through_c_abi():
    from b200qp import lowlevel
    s = lowlevel.default_settings()
    assert s.linsys_solver == 2 and s.rho == 0.1 and s.sigma == 1e-6 and s.alpha == 1.6
    assert s.max_iter == 4000 and s.check_termination == 25 and s.cg_max_iter == 100 and s.cg_tol_fraction == 0.1
    assert lowlevel.lib().b200qp_capabilities() == 0x02 | 0x08 | 0x10
    assert lowlevel.variants() == ["group", "stream-tma", "scalar", "tile-tma"]


def test_extension_surface_matches_reference_facade():
    import osqp_cuda as ext
    assert ext.OSQP_USE_FLOAT == 0 and ext.OSQP_USE_LONG == 0 and ext.OSQP_INFTY == 1e30
    st = list(ext.osqp_status_type.__members__)
    assert st == ["OSQP_SOLVED", "OSQP_SOLVED_INACCURATE", "OSQP_PRIMAL_INFEASIBLE",
                  "OSQP_PRIMAL_INFEASIBLE_INACCURATE", "OSQP_DUAL_INFEASIBLE", "OSQP_DUAL_INFEASIBLE_INACCURATE",
                  "OSQP_MAX_ITER_REACHED", "OSQP_TIME_LIMIT_REACHED", "OSQP_NON_CVX", "OSQP_SIGINT", "OSQP_UNSOLVED"]
    assert [int(v.value) for v in ext.osqp_status_type.__members__.values()] == list(range(1, 12))
    assert len(ext.osqp_error_type.__members__) == 11
    settings_fields = [k for k in vars(ext.OSQPSettings) if not k.startswith("_")]
    for f in ("device linsys_solver verbose warm_starting scaling polishing rho rho_is_vec sigma alpha cg_max_iter "
              "cg_tol_reduction cg_tol_fraction cg_precond adaptive_rho adaptive_rho_interval adaptive_rho_fraction "
              "adaptive_rho_tolerance max_iter eps_abs eps_rel eps_prim_inf eps_dual_inf scaled_termination "
              "check_termination check_dualgap time_limit delta polish_refine_iter").split():
        assert f in settings_fields
    assert len(settings_fields) == 29
    info_fields = [k for k in vars(ext.OSQPInfo) if not k.startswith("_")]
    for f in ("status status_val status_polish obj_val dual_obj_val prim_res dual_res duality_gap iter rho_updates "
              "rho_estimate setup_time solve_time update_time polish_time run_time primdual_int rel_kkt_error").split():
        assert f in info_fields
    for m in ("solve", "warm_start", "update_data_vec", "update_data_mat", "update_settings", "update_rho",
              "get_settings", "solution", "info"):
        assert hasattr(ext.OSQPSolver, m)
    M = sparse.random(5, 4, density=0.5, format="csc", random_state=0)
    c = ext.CSC(M)
    assert (c.m, c.n, c.nzmax, c.nz) == (5, 4, M.nnz, -1)
    np.testing.assert_array_equal(c.p, M.indptr)
    s = ext.OSQPSettings()
    s.linsys_solver = ext.osqp_linsys_solver_type.OSQP_DIRECT_SOLVER
    assert s.linsys_solver == ext.OSQP_DIRECT_SOLVER
    assert ext.OSQP_DIAGONAL_PRECONDITIONER == ext.osqp_precond_type.OSQP_DIAGONAL_PRECONDITIONER


def test_facade_names():
    import osqp
    assert osqp.algebra_available("cuda") and not osqp.algebra_available("builtin")
    assert osqp.default_algebra() == "cuda"
    assert osqp.SolverStatus.OSQP_SOLVED == 1 and osqp.SolverStatus.OSQP_UNSOLVED == 11
    assert osqp.SolverError.OSQP_NONCVX_ERROR == 4
    assert np.isnan(osqp.constant("OSQP_NAN")) and osqp.constant("OSQP_INFTY") == 1e30
    assert osqp.OSQPException(4) == osqp.SolverError.OSQP_NONCVX_ERROR
    with pytest.raises(RuntimeError):
        osqp.OSQP(algebra="builtin")
    import osqp.cuda
    assert osqp.cuda.OSQP().algebra == "cuda"
    with pytest.raises(TypeError):
        osqp.OSQP().setup(np.eye(2), np.zeros(2), sparse.eye(2, format="csc"), np.zeros(2), np.ones(2))


def test_no_cpu_fallback_without_device():
    """With no CUDA device the engine must raise -- never compute on the host."""
    import osqp
    import osqp_cuda
    if osqp_cuda.device_count() > 0:
        pytest.skip("a CUDA device is present")
    P = sparse.diags([11.0, 0.0], format="csc")
    A = sparse.csc_matrix([[-1, 0], [0, -1], [-1, -3], [2, 5], [3, 4]], dtype=float)
    prob = osqp.OSQP()
    with pytest.raises(osqp.OSQPException) as ei:
        prob.setup(P, np.array([3.0, 4.0]), A, -1e6 * np.ones(5), np.array([0.0, 0.0, -15, 100, 80]), verbose=False)
    assert ei.value == osqp.SolverError.OSQP_ALGEBRA_LOAD_ERROR
    from b200qp import lowlevel
    with pytest.raises(RuntimeError):
        lowlevel.spmv_host(sparse.eye(4, format="csr"), np.ones(4))


def test_batch_front_end_without_device_and_shape_checks():
    """osqp.OSQPBatch: argument shapes are checked on the host; without a CUDA device setup raises (no fallback)."""
    import osqp
    import osqp_cuda
    P = sparse.diags([1.0, 2.0, 3.0], format="csc")
    A = sparse.eye(3, format="csc")
    q = np.zeros((4, 3))
    l, u = -np.ones((4, 3)), np.ones((4, 3))
    prob = osqp.OSQPBatch()
    with pytest.raises(AssertionError):
        prob.setup(P, np.zeros((4, 2)), A, l, u, verbose=False)               # q has the wrong width
    with pytest.raises(AssertionError):
        prob.setup(P, q, A, l[:3], u, verbose=False)                          # l has the wrong batch size
    for name in ("OSQPBatchSolver",):
        assert hasattr(osqp_cuda, name)
    for m in ("solve", "update_data_vec", "update_data_mat", "warm_start", "update_settings", "update_rho",
              "get_solution", "get_info", "get_stats"):
        assert hasattr(osqp_cuda.OSQPBatchSolver, m)
    for m in ("adjoint_derivative_compute", "adjoint_derivative_get_mat", "adjoint_derivative_get_vec"):
        assert hasattr(osqp_cuda.OSQPSolver, m) and hasattr(osqp.OSQP, m)
    if osqp_cuda.device_count() == 0:
        with pytest.raises(osqp.OSQPException) as ei:
            prob.setup(P, q, A, l, u, verbose=False)
        assert ei.value == osqp.SolverError.OSQP_ALGEBRA_LOAD_ERROR


def test_setup_validation_error_codes():
    """Data / settings validation happens before any device work, so it is testable on CPU."""
    import osqp
    P = sparse.diags([1.0, 1.0], format="csc")
    A = sparse.eye(2, format="csc")
    prob = osqp.OSQP()
    with pytest.raises(osqp.OSQPException) as ei:
        prob.setup(P, np.zeros(2), A, np.ones(2), np.zeros(2), verbose=False)      # l > u
    assert ei.value == osqp.SolverError.OSQP_DATA_VALIDATION_ERROR
    with pytest.raises(osqp.OSQPException) as ei:
        prob.setup(P, np.zeros(2), A, np.zeros(2), np.ones(2), verbose=False, eps_abs=-1.0)
    assert ei.value == osqp.SolverError.OSQP_SETTINGS_VALIDATION_ERROR
    with pytest.raises(osqp.OSQPException) as ei:
        prob.setup(P, np.zeros(2), A, np.zeros(2), np.ones(2), verbose=False, solver_type="direct")
    assert ei.value == osqp.SolverError.OSQP_LINSYS_SOLVER_INIT_ERROR
    with pytest.raises(ValueError):
        prob.setup(P, np.zeros(2), A, np.zeros(2), np.ones(2), foo=42)


def test_problem_generators():
    from b200qp import problems
    P, q, A, l, u = problems.random_qp(200, 300, nnz_per_row=10, seed=3)
    assert A.shape == (300, 200) and A.nnz == 3000 and np.all(np.diff(A.tocsr().indptr) == 10)
    assert abs(P - P.T).max() < 1e-15 and np.all(l <= u)
    w = np.linalg.eigvalsh(P.toarray())
    assert w.min() > 0
    P, q, A, l, u = problems.lasso_qp(20, 40, density=0.2)
    assert P.shape == (80, 80) and A.shape == (40 + 40 + 80, 80)
    P, q, A, l, u = problems.svm_qp(10, 40, density=0.2, box_rows=5)
    assert P.shape == (50, 50) and A.shape == (85, 50)
    P, q, A, l, u = problems.mpc_qp(seed=7)
    assert P.shape == (120, 120) and A.shape == (240, 120)
