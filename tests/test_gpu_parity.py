"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes
osqp.OSQP -> osqp_cuda (pybind11) -> libb200qp.so C ABI -> the persistent sm_100a ADMM kernel.

Checked against, on identical (P, q, A, l, u):
  * the reference's golden vectors (tests/golden/cases/*.npz `ref_*`) at the reference's own
    *builtin/direct* tolerances atol 1e-3, rtol 1e-4, 4 decimals (conftest.py:15-17) -- stricter
    than what the reference asks of its CUDA algebra (1e-2, 1e-3, 2 decimals; conftest.py:26-29);
  * recorded outputs of the reference's pure-Python solver (`purepy_*`);
  * the CPU oracle run live (oracle/osqp_oracle.py), direct and engine-spec indirect mode;
  * the behavioural pins of the reference tests (iteration-count equalities, statuses).
Tolerances on x*, y*, obj for tight-eps cases: 1e-6 relative (BASELINE.json north_star).
"""
import numpy as np
import numpy.testing as npt
import pytest
from scipy import sparse

from case_runner import case_names, load_case, run_case_engine, run_case_oracle, status_name

pytestmark = pytest.mark.gpu

ATOL, RTOL, DEC = 1e-3, 1e-4, 4
CASES = case_names()
REL = 1e-6


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1.0, float(np.max(np.abs(b))))) if len(b) else 0.0


@pytest.mark.parametrize("name", CASES)
def test_engine_vs_reference_golden(name):
    case = load_case(name)
    e = case["expect"]
    if not any(k.startswith("ref_") for k in e):
        pytest.skip("no reference golden for this case")
    res, _ = run_case_engine(case)
    if "ref_x_val" in e:
        assert status_name(res) == "solved"
        npt.assert_allclose(res.x, e["ref_x_val"], rtol=RTOL, atol=ATOL)
        if case["m"]:
            npt.assert_allclose(res.y, e["ref_y_val"], rtol=RTOL, atol=ATOL)
        npt.assert_almost_equal(res.info.obj_val, float(e["ref_obj"]), decimal=DEC)
    if "ref_normalized_prim_inf_cert_correct" in e:
        assert status_name(res) == "primal infeasible"
        v = res.prim_inf_cert / np.linalg.norm(res.prim_inf_cert)
        # golden entries are O(1) for the violated pair and O(1e-5) elsewhere; an inexact inner solve
        # perturbs the small ones at the 1e-5 level (see the oracle's indirect mode)
        npt.assert_allclose(v, e["ref_normalized_prim_inf_cert_correct"], atol=1e-4)
    if "ref_normalized_dual_inf_cert" in e:
        assert status_name(res) == "dual infeasible"
        v = res.dual_inf_cert / np.linalg.norm(res.dual_inf_cert)
        assert np.allclose(v, e["ref_normalized_dual_inf_cert"], atol=1e-6)


@pytest.mark.parametrize("name", CASES)
def test_engine_vs_recorded_purepy(name):
    case = load_case(name)
    e = case["expect"]
    if "purepy_status" not in e:
        pytest.skip("purepy cannot run this case (m == 0)")
    res, _ = run_case_engine(case, max_iter=40000)
    want = str(e["purepy_status"])
    if name in ("test_primal_and_dual_infeasible", "test_primal_and_dual_infeasible_ws"):
        assert status_name(res) in ("primal infeasible", "dual infeasible")   # reference accepts either
        return
    assert status_name(res) == want
    if want != "solved":
        return
    eps = case["settings"].get("eps_abs", 1e-3)
    if case["settings"].get("polishing"):
        tol = 1e-5 if res.info.status_polish == 1 else 10 * eps
    else:
        tol = REL if eps <= 1e-8 else 20 * eps
    if name == "test_feasibility_problem":
        tol = 5e-4     # P = 0, square A: x* = A^-1 u is ill-conditioned; reference itself excuses indirect here
    assert _rel(res.x, e["purepy_x"]) <= tol
    assert _rel(res.y, e["purepy_y"]) <= tol * (10 if eps > 1e-8 else 1)
    assert abs(res.info.obj_val - float(e["purepy_obj"])) <= tol * max(1.0, abs(float(e["purepy_obj"])))


@pytest.mark.parametrize("name", ["test_basic_QP", "test_solve", "polish_random_tight", "test_warm_start",
                                  "c1_random_qp", "test_update_P_A_allind", "test_update_bounds"])
def test_engine_vs_live_oracle(name):
    """Engine vs the oracle's engine-spec (indirect) mode: same algorithm, so statuses agree and
    iteration counts agree up to floating-point summation order; vs direct mode: same optimum."""
    case = load_case(name)
    kw = dict(max_iter=40000, check_termination=1, adaptive_rho_interval=50)
    res, prob = run_case_engine(case, **kw)
    ind = run_case_oracle(case, "indirect", cg_max_iter=100, cg_tol_fraction=0.1, rho_eq_factor=10.0, **kw)
    dire = run_case_oracle(case, "direct", **kw)
    assert status_name(res) == ind.info.status == dire.info.status == "solved"
    # identical algorithm, different floating-point summation order: short runs agree to a few iterations,
    # long ill-conditioned ones (c1: thousands of iterations at eps 1e-9) drift by a few percent
    assert abs(res.info.iter - ind.info.iter) <= max(3, 0.2 * ind.info.iter)
    for ref in (ind, dire):
        assert _rel(res.x, ref.x) <= REL
        assert _rel(res.y, ref.y) <= 10 * REL
        assert abs(res.info.obj_val - ref.info.obj_val) <= REL * max(1.0, abs(ref.info.obj_val))
    st = prob.stats()
    assert st["admm_iters"] == res.info.iter and st["cg_iters"] > 0 and st["kernel_ms"] > 0


# ------------------------------------------------------------------ behavioural pins of the reference tests
def _basic(**over):
    import osqp
    case = load_case("test_basic_QP")
    st = dict(case["settings"], verbose=False)
    st.update(over)
    prob = osqp.OSQP(algebra="cuda")
    prob.setup(P=case["P"], q=case["q"], A=case["A"], l=case["l"], u=case["u"], **st)
    return prob, case


def test_update_max_iter():                      # basic_test.py:102-106
    prob, _ = _basic()
    prob.update_settings(max_iter=80)
    assert prob.solve().info.status_val == prob.constant("OSQP_MAX_ITER_REACHED")


def test_update_check_termination():             # basic_test.py:109-113
    prob, case = _basic()
    prob.update_settings(check_termination=0)
    assert prob.solve().info.iter == case["settings"]["max_iter"]


def test_update_rho_same_iterations():           # basic_test.py:116-128
    prob, case = _basic()
    it_default = prob.solve().info.iter
    other, _ = _basic(rho=0.7)
    other.update_settings(rho=case["settings"]["rho"])
    assert other.solve().info.iter == it_default


def test_upper_triangular_P():                   # basic_test.py:131-148
    import osqp
    case = load_case("test_solve")
    full = case["P"] + sparse.triu(case["P"], 1).T
    out = []
    for Pm in (full.tocsc(), sparse.triu(full, format="csc")):
        prob = osqp.OSQP()
        prob.setup(P=Pm, q=case["q"], A=case["A"], l=case["l"], u=case["u"], verbose=False, eps_abs=1e-8, eps_rel=1e-8)
        out.append(prob.solve())
    npt.assert_allclose(out[0].x, out[1].x, rtol=1e-12, atol=1e-12)
    npt.assert_allclose(out[0].y, out[1].y, rtol=1e-12, atol=1e-12)
    assert out[0].info.iter == out[1].info.iter


def test_update_invalid_setting():               # basic_test.py:151-154
    prob, _ = _basic()
    with pytest.raises(ValueError):
        prob.update_settings(foo=42)


def test_update_bounds_validation():
    import osqp
    prob, _ = _basic()
    rc = prob._solver.update_data_vec(q=None, l=np.ones(5), u=np.zeros(5))
    assert rc == osqp.SolverError.OSQP_DATA_VALIDATION_ERROR


@pytest.mark.parametrize("eps", [1e-8, 1e-2])
def test_warm_start_pins(eps):                   # warm_start_test.py:25-57 (direct and indirect tolerances)
    import osqp
    case = load_case("test_warm_start")
    prob = osqp.OSQP()
    prob.setup(P=case["P"], q=case["q"], A=case["A"], l=case["l"], u=case["u"], verbose=False, adaptive_rho=False,
               eps_abs=eps, eps_rel=eps, polishing=False, check_termination=1)
    res = prob.solve()
    x_opt, y_opt, tot = res.x, res.y, res.info.iter
    prob.warm_start(x=np.zeros(case["n"]), y=np.zeros(case["m"]))
    assert prob.solve().info.iter == tot
    prob.warm_start(x=x_opt, y=y_opt)
    assert prob.solve().info.iter < 10


def test_non_convex():                           # non_convex_test.py:22-60
    import osqp
    P = sparse.triu([[2.0, 5.0], [5.0, 1.0]], format="csc")
    q = np.array([3, 4])
    A = sparse.csc_matrix([[-1.0, 0.0], [0.0, -1.0], [-1.0, 3.0], [2.0, 5.0], [3.0, 4]])
    u = np.array([0.0, 0.0, -15, 100, 80])
    l = -np.inf * np.ones(len(u))
    prob = osqp.OSQP()
    prob.setup(P=P, q=q, A=A, l=l, u=u, solver_type="indirect", sigma=1e-6, verbose=False)
    res = prob.solve()
    assert res.info.status_val in (prob.constant("OSQP_MAX_ITER_REACHED"), prob.constant("OSQP_NON_CVX"))
    prob = osqp.OSQP()
    prob.setup(P=P, q=q, A=A, l=l, u=u, sigma=5, verbose=False)
    res = prob.solve()
    assert res.info.status_val == prob.constant("OSQP_NON_CVX")
    assert np.isnan(res.info.obj_val)
    with pytest.raises(osqp.OSQPException):
        prob.solve(raise_error=True)


def test_time_limit_and_verbose(capfd):
    import osqp
    from b200qp import problems
    P, q, A, l, u = problems.random_qp(2000, 3000, seed=11)
    prob = osqp.OSQP()
    prob.setup(P, q, A, l, u, verbose=True, eps_abs=1e-12, eps_rel=1e-12, max_iter=200000, time_limit=0.05)
    res = prob.solve()
    assert res.info.status_val == prob.constant("OSQP_TIME_LIMIT_REACHED")
    out = capfd.readouterr().out
    assert "iter" in out and "status:" in out


def test_multithread_distinct_solvers():         # multithread_test.py:12-55 (GIL released in solve)
    import osqp
    from multiprocessing.pool import ThreadPool
    from b200qp import problems
    data = [problems.random_qp(300, 500, seed=s) for s in range(8)]

    def f(i):
        P, q, A, l, u = data[i]
        m = osqp.OSQP()
        m.setup(P, q, A, l, u, verbose=False, eps_abs=1e-6, eps_rel=1e-6, max_iter=20000)
        r = m.solve()
        return r.info.status, r.info.obj_val

    serial = [f(i) for i in range(8)]
    par = ThreadPool(4).map(f, range(8))
    assert all(s[0] == "solved" for s in serial)
    assert serial == par     # bitwise reproducible and thread-safe


# ------------------------------------------------------------------ BASELINE configs at reduced and full size
def _solve_and_check(P, q, A, l, u, eps=1e-6, **st):
    import osqp
    from b200qp import problems
    prob = osqp.OSQP()
    prob.setup(P, q, A, l, u, verbose=False, eps_abs=eps, eps_rel=eps, **st)
    res = prob.solve()
    rep = problems.kkt_report(P, q, A, l, u, res.x, res.y)
    return res, rep, prob


def test_c2_lasso_small_vs_oracle():
    from b200qp import problems
    from oracle.osqp_oracle import solve_qp
    P, q, A, l, u = problems.lasso_qp(60, 120, density=0.15, seed=1)
    res, rep, _ = _solve_and_check(P, q, A, l, u, eps=1e-9, max_iter=100000, check_termination=1)
    ora = solve_qp(P, q, A, l, u, eps_abs=1e-9, eps_rel=1e-9, max_iter=100000, check_termination=1)
    assert res.info.status == ora.info.status == "solved"
    assert _rel(res.x, ora.x) <= REL and abs(res.info.obj_val - ora.info.obj_val) <= REL * max(1, abs(ora.info.obj_val))


def test_c3_svm_small_warm_start_polish_vs_oracle():
    """Config 3 flow: adaptive rho, solve, perturb q, warm start from the previous solution, polish."""
    import osqp
    from b200qp import problems
    from oracle.osqp_oracle import OracleSolver
    P, q, A, l, u = problems.svm_qp(40, 160, density=0.3, seed=2, box_rows=10)
    kw = dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=100000, adaptive_rho=1)
    prob = osqp.OSQP()
    prob.setup(P, q, A, l, u, verbose=False, **kw)
    r1 = prob.solve()
    ora = OracleSolver(P, q, A, l, u, **kw)
    o1 = ora.solve()
    assert r1.info.status == o1.info.status == "solved"
    assert _rel(r1.x, o1.x) <= REL
    q2 = q * (1 + 0.01 * np.cos(np.arange(len(q))))
    prob.update(q=q2)
    prob.warm_start(x=r1.x, y=r1.y)
    r2 = prob.solve()
    ora.update(q=q2)
    ora.warm_start(x=o1.x, y=o1.y)
    o2 = ora.solve()
    assert r2.info.status == "solved" and r2.info.iter < r1.info.iter
    assert _rel(r2.x, o2.x) <= REL and abs(r2.info.obj_val - o2.info.obj_val) <= REL * max(1, abs(o2.info.obj_val))


def test_c1_sized_random_qp_kkt():
    from b200qp import problems
    P, q, A, l, u = problems.random_qp(100, 200, nnz_per_row=3, seed=0, eq_frac=0.15, onesided_frac=0.1)
    res, rep, _ = _solve_and_check(P, q, A, l, u, eps=1e-9, max_iter=100000)
    assert res.info.status == "solved"
    assert rep["prim"] < 1e-6 and rep["dual"] < 1e-6 and rep["comp"] < 1e-5


def test_c4_full_size_properties():
    """BASELINE config 4 at full size (n = m = 1e6, nnz(A) = 1e7): the oracle cannot finish, so check
    size-independent properties: KKT residuals of the returned point, determinism, iteration accounting."""
    from b200qp import problems
    P, q, A, l, u = problems.random_qp(1_000_000, 1_000_000, nnz_per_row=10, seed=3)
    res, rep, prob = _solve_and_check(P, q, A, l, u, eps=1e-4, max_iter=2000)
    assert res.info.status == "solved"
    scale = max(1.0, float(np.max(np.abs(q))))
    assert rep["prim"] <= 1e-3 and rep["dual"] <= 1e-3 * scale
    st = prob.stats()
    assert st["admm_iters"] == res.info.iter and st["nnzA"] == 10_000_000
    again = prob.solve()                      # warm-started from the solution: immediate
    assert again.info.status == "solved" and again.info.iter <= 25


def test_c2_lasso_full_size_properties():
    """BASELINE config 2 at full size: Lasso-as-QP n = 1e5, m = 2e5, data block 5e4 x 2.5e4 at 5 % density
    (6.25e7 nnz, rows of 1250 and columns of 2500 entries).  Oracle cannot finish: KKT residuals instead."""
    from b200qp import problems
    P, q, A, l, u = problems.lasso_qp(25_000, 50_000, density=0.05, seed=1, box=True)
    assert P.shape[0] == 100_000 and A.shape[0] == 200_000
    res, rep, prob = _solve_and_check(P, q, A, l, u, eps=1e-3, max_iter=4000)
    assert res.info.status in ("solved", "solved inaccurate")
    scale = max(1.0, float(np.max(np.abs(q))))
    assert rep["prim"] <= 5e-2 and rep["dual"] <= 5e-2 * scale
    print("C2 full:", res.info.status, res.info.iter, prob.stats()["cg_iters"], prob.stats()["kernel_ms"], rep)


def test_c3_svm_full_size_warm_start_polish():
    """BASELINE config 3 at full size: SVM QP n = 5e4, m = 1e5 (n_feat 1e4, n_data 4e4, +2e4 box rows), adaptive rho,
    then perturbed q + warm start + polish.  At eps 1e-5 the active set is identified and the polish must SUCCEED, cold
    and warm-started, leaving a point at KKT residuals ~1e-12 (a factorisation-class answer from an iterative engine);
    at the default eps 1e-3 the reference algorithm's own acceptance test may reject the guess (-1), as it does on the
    reduced instances of tests/test_gpu_parity2.py::test_polish_status_follows_the_reference_algorithm."""
    import osqp
    from b200qp import problems
    P, q, A, l, u = problems.svm_qp(10_000, 40_000, density=0.01, seed=2, box_rows=20_000)
    assert P.shape[0] == 50_000 and A.shape[0] == 100_000
    prob = osqp.OSQP()
    prob.setup(P, q, A, l, u, verbose=False, adaptive_rho=True, polishing=True, max_iter=20000, eps_abs=1e-5, eps_rel=1e-5)
    r1 = prob.solve()
    assert r1.info.status == "solved" and r1.info.status_polish == 1
    rep = problems.kkt_report(P, q, A, l, u, r1.x, r1.y)
    assert rep["prim"] <= 1e-9 and rep["dual"] <= 1e-9
    q2 = q * (1 + 0.01 * np.cos(np.arange(len(q))))
    prob.update(q=q2)
    prob.warm_start(x=r1.x, y=r1.y)
    r2 = prob.solve()
    assert r2.info.status == "solved" and r2.info.iter <= r1.info.iter
    assert r2.info.status_polish == 1
    rep2 = problems.kkt_report(P, q2, A, l, u, r2.x, r2.y)
    assert rep2["prim"] <= 1e-9 and rep2["dual"] <= 1e-9
    # default tolerances: still a solve within the loose bounds, polish status whatever the acceptance test says
    loose = osqp.OSQP()
    loose.setup(P, q, A, l, u, verbose=False, adaptive_rho=True, polishing=True, max_iter=10000)
    r3 = loose.solve()
    assert r3.info.status == "solved" and r3.info.status_polish in (1, -1)
    rep3 = problems.kkt_report(P, q, A, l, u, r3.x, r3.y)
    assert rep3["prim"] <= 1e-2 and rep3["dual"] <= 1e-2
    print("C3 full:", r1.info.iter, r2.info.iter, r1.info.status_polish, r2.info.status_polish, r3.info.status_polish, rep, rep2)
