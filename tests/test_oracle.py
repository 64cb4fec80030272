"""Pin the CPU oracle (oracle/osqp_oracle.py, direct mode) against
  (a) the reference's own golden vectors (src/osqp/tests/solutions/*.npz, copied as data into the
      case files by tests/golden/make_golden.py) at the reference's builtin tolerances
      (conftest.py:15-17: atol 1e-3, rtol 1e-4, 4 decimals), and
  (b) the output of the reference's in-tree solver osqppurepy recorded in the build container:
      same status, same iteration count, x/y/obj to 1e-9.
Also checks the engine-spec indirect mode converges to the same optimum.  CPU only."""
import numpy as np
import numpy.testing as npt
import pytest

from case_runner import case_names, load_case, run_case_oracle

ATOL, RTOL, DEC = 1e-3, 1e-4, 4
CASES = case_names()


def _like_purepy(case):
    """purepy checks termination every iteration and adapts rho every 200 (ref:159,164)."""
    st = case["settings"]
    return dict(check_termination=st.get("check_termination", 1) or 1,
                adaptive_rho_interval=st.get("adaptive_rho_interval", 200))


def test_case_inventory():
    assert len(CASES) >= 28
    ref = [c for c in CASES if any(k.startswith("ref_") for k in load_case(c)["expect"])]
    assert len(ref) >= 22  # every golden file the reference's tests use on this path


@pytest.mark.parametrize("name", CASES)
def test_direct_vs_reference_golden(name):
    case = load_case(name)
    e = case["expect"]
    if not any(k.startswith("ref_") for k in e):
        pytest.skip("no reference golden for this case")
    # Certificates: the golden was produced by the C solver at its default cadence
    # (check_termination=25); checking at iteration 1 (purepy's cadence) detects earlier, before the
    # O(1e-5) entries of delta_y have settled.  Solutions: cadence is irrelevant at these tolerances.
    kw = {} if "ref_normalized_prim_inf_cert_correct" in e else _like_purepy(case)
    res = run_case_oracle(case, "direct", **kw)
    if "ref_x_val" in e:
        npt.assert_allclose(res.x, e["ref_x_val"], rtol=RTOL, atol=ATOL)
        if case["m"]:
            npt.assert_allclose(res.y, e["ref_y_val"], rtol=RTOL, atol=ATOL)
        npt.assert_almost_equal(res.info.obj_val, float(e["ref_obj"]), decimal=DEC)
    if "ref_normalized_prim_inf_cert_correct" in e:
        assert res.info.status == "primal infeasible"
        v = res.prim_inf_cert / np.linalg.norm(res.prim_inf_cert)
        assert np.allclose(v, e["ref_normalized_prim_inf_cert_correct"])
    if "ref_normalized_dual_inf_cert" in e:
        assert res.info.status == "dual infeasible"
        v = res.dual_inf_cert / np.linalg.norm(res.dual_inf_cert)
        assert np.allclose(v, e["ref_normalized_dual_inf_cert"])


@pytest.mark.parametrize("name", CASES)
def test_direct_vs_purepy(name):
    case = load_case(name)
    e = case["expect"]
    if "purepy_status" not in e:
        pytest.skip("purepy cannot run this case (m == 0)")
    res = run_case_oracle(case, "direct", **_like_purepy(case))
    assert res.info.status == str(e["purepy_status"])
    assert res.info.iter == int(e["purepy_iter"])
    if "purepy_x" in e and res.info.status == "solved":
        npt.assert_allclose(res.x, e["purepy_x"], rtol=1e-9, atol=1e-9)
        npt.assert_allclose(res.y, e["purepy_y"], rtol=1e-9, atol=1e-9)
        assert abs(res.info.obj_val - float(e["purepy_obj"])) <= 1e-9 * max(1, abs(float(e["purepy_obj"])))
        assert res.info.status_polish == int(e["purepy_status_polish"])
    if "purepy_prim_inf_cert" in e:
        a, b = res.prim_inf_cert, e["purepy_prim_inf_cert"]
        npt.assert_allclose(a / np.linalg.norm(a), b / np.linalg.norm(b), atol=1e-9)
    if "purepy_dual_inf_cert" in e:
        a, b = res.dual_inf_cert, e["purepy_dual_inf_cert"]
        npt.assert_allclose(a / np.linalg.norm(a), b / np.linalg.norm(b), atol=1e-9)


@pytest.mark.parametrize("name", ["test_basic_QP", "test_solve", "polish_random_tight", "test_warm_start",
                                  "test_update_P_A_allind", "c1_random_qp"])
def test_indirect_spec_reaches_same_optimum(name):
    """The engine-spec PCG mode is an inexact-ADMM variant: different trajectory, same optimum."""
    case = load_case(name)
    e = case["expect"]
    res = run_case_oracle(case, "indirect", cg_max_iter=100, max_iter=40000, **_like_purepy(case))
    assert res.info.status == "solved"
    scale = max(1.0, np.max(np.abs(e["purepy_x"])))
    assert np.max(np.abs(res.x - e["purepy_x"])) <= 1e-6 * scale
    assert abs(res.info.obj_val - float(e["purepy_obj"])) <= 1e-6 * max(1.0, abs(float(e["purepy_obj"])))


def test_warm_start_pins():
    """warm_start_test.py:49-57: zero warm start => identical iteration count; optimal => < 10."""
    from oracle.osqp_oracle import OracleSolver
    case = load_case("test_warm_start")
    for mode in ("direct", "indirect"):
        s = OracleSolver(case["P"], case["q"], case["A"], case["l"], case["u"], linsys=mode,
                         cg_max_iter=100, **case["settings"])
        r1 = s.solve()
        it1, x, y = r1.info.iter, r1.x.copy(), r1.y.copy()
        s.warm_start(x=np.zeros(case["n"]), y=np.zeros(case["m"]))
        assert s.solve().info.iter == it1
        s.warm_start(x=x, y=y)
        assert s.solve().info.iter < 10


def test_update_rho_pin():
    """basic_test.py:116-128: setup(rho=.7) + update_rho(.1) == setup(rho=.1) in iterations."""
    from oracle.osqp_oracle import OracleSolver
    case = load_case("test_basic_QP")
    a = OracleSolver(case["P"], case["q"], case["A"], case["l"], case["u"], **case["settings"]).solve()
    st = dict(case["settings"], rho=0.7)
    b = OracleSolver(case["P"], case["q"], case["A"], case["l"], case["u"], **st)
    b.update_rho(0.1)
    assert b.solve().info.iter == a.info.iter


def test_oracle_adjoint_derivative_matches_finite_differences():
    """The checker of the engine's adjoint derivatives (tests/test_gpu_derivative.py) against finite differences of
    complete oracle solves; problem family of the reference's derivative_test.py:32-47."""
    import numpy.random as npr
    from scipy import sparse
    from scipy.optimize import approx_fprime
    from oracle.osqp_oracle import adjoint_derivative, solve_qp
    npr.seed(1)
    n, m = 8, 5
    L = npr.randn(n, n - 1)
    P = sparse.csc_matrix(L.dot(L.T) + 0.1 * sparse.eye(n))
    x0, s0 = npr.randn(n), npr.rand(m)
    A = sparse.csc_matrix(npr.randn(m, n))
    u, l = A.dot(x0) + s0, A.dot(x0) - s0
    q, tx, ty = npr.randn(n), npr.randn(n), npr.randn(m)
    kw = dict(eps_abs=1e-10, eps_rel=1e-10, max_iter=200000)

    def loss(P_, q_, A_, l_, u_):
        r = solve_qp(P_, q_, A_, l_, u_, linsys="direct", **kw)
        assert r.info.status == "solved"
        return 0.5 * np.sum((r.x - tx) ** 2) + 0.5 * np.sum((r.y - ty) ** 2), r

    _, r = loss(P, q, A, l, u)
    dP, dq, dA, dl, du = adjoint_derivative(P, q, A, l, u, r.x, r.y, r.x - tx, r.y - ty)
    fd = approx_fprime(q, lambda v: loss(P, v, A, l, u)[0], 1e-6)
    np.testing.assert_allclose(dq, fd, atol=1e-4)
    fd = approx_fprime(l, lambda v: loss(P, q, A, v, u)[0], 1e-6)
    np.testing.assert_allclose(dl, fd, atol=1e-4)
    fd = approx_fprime(u, lambda v: loss(P, q, A, l, v)[0], 1e-6)
    np.testing.assert_allclose(du, fd, atol=1e-4)
    Ad = A.toarray()
    fd = approx_fprime(Ad.ravel(), lambda v: loss(P, q, sparse.csc_matrix(v.reshape(m, n)), l, u)[0], 1e-6).reshape(m, n)
    np.testing.assert_allclose(dA, fd, atol=1e-3)


def test_batch_element_semantics_setup_then_update_equals_fresh_setup_when_q_is_shared():
    """The batched engine defines element b as setup(element 0) + update(l[b], u[b]) + solve().  For config 5 (MPC:
    shared P, q, A; per-element bounds) that is the same solve as a fresh setup on element b, because the
    equilibration depends on P, A and q only (ref:389-497) -- iteration for iteration."""
    from b200qp import problems            # tests/conftest.py puts osqp-python_b200/ on sys.path
    from oracle.osqp_oracle import OracleSolver
    P, q, A, l, u = problems.mpc_batch(4)
    for b in (1, 3):
        s = OracleSolver(P, q, A, l[0], u[0], linsys="direct")
        s.update(l=l[b], u=u[b])
        r1 = s.solve()
        r2 = OracleSolver(P, q, A, l[b], u[b], linsys="direct").solve()
        assert r1.info.status == r2.info.status == "solved" and r1.info.iter == r2.info.iter
        np.testing.assert_allclose(r1.x, r2.x, rtol=0, atol=1e-12)
        np.testing.assert_allclose(r1.y, r2.y, rtol=0, atol=1e-12)
