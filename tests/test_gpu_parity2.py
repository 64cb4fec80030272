"""Round-2 GPU parity tests (pytest -m gpu): what VERDICT r1 listed as missing.

  * the `*_INACCURATE` statuses (reference: src/osqppurepy/_osqp.py:1018-1022, 1263-1266), produced on purpose and
    compared with the oracle's engine-spec mode at the same iteration budget;
  * check_dualgap / dual_obj_val / duality_gap (bindings.cpp.in:479-492) against the oracle and strong duality;
  * polish against the oracle's factorisation-based polish (same status_polish; polished x, y within 1e-7);
  * time_limit with check_termination = 0, partial warm starts, NaN solution on NON_CVX;
  * BASELINE configs 2, 3, 4 AT THE NAMED SIZE against the oracle: the first ADMM iterates with near-exact inner
    solves (both sides solve every KKT system to 1e-9, so the iterates are the same mathematical sequence and must
    agree to rounding), and for config 4 the complete default-tolerance solve (status, iteration count, x, objective).
Everything goes osqp.OSQP -> osqp_cuda (pybind11) -> C ABI -> persistent sm_100a kernel.
"""
import time

import numpy as np
import pytest
from scipy import sparse

from case_runner import load_case

pytestmark = pytest.mark.gpu

ENGINE_SPEC = dict(cg_max_iter=100, cg_tol_fraction=0.1, rho_eq_factor=10.0)   # the engine's defaults, for the oracle


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b))))) if b.size else 0.0


def _engine(P, q, A, l, u, **st):
    import osqp
    prob = osqp.OSQP(algebra="cuda")
    prob.setup(P=P, q=q, A=A, l=l, u=u, verbose=False, **st)
    return prob


def _oracle(P, q, A, l, u, linsys="indirect", **st):
    from oracle.osqp_oracle import OracleSolver
    kw = dict(ENGINE_SPEC) if linsys == "indirect" else {}
    kw.update(st)
    return OracleSolver(P, q, A, l, u, linsys=linsys, **kw)


# ------------------------------------------------------------------------------------------------ *_INACCURATE
def _find_budget(make_oracle, want, lo=1, hi=400):
    """Iteration budgets (max_iter) at which the oracle, whose termination is evaluated only at max_iter, ends in
    `want`: the longest contiguous run of them (the residuals of an inexact-PCG ADMM are not monotone, and the dual
    infeasibility test of the two-variable LP alternates with the parity of the iteration)."""
    hits = [k for k in range(lo, hi) if make_oracle(k).solve().info.status_val == want]
    best, run = [], []
    for k in hits:
        run = run + [k] if run and k == run[-1] + 1 else [k]
        if len(run) > len(best):
            best = run
    return best


def _pinf_problem(seed=5):
    """A small QP whose last two rows contradict each other (a'x <= -1 and a'x >= 1) while the others couple all
    variables: the infeasibility certificate emerges gradually (the reference's own fixture is detected at iteration 1,
    so it never passes through the 10x window)."""
    rng = np.random.default_rng(seed)
    n, m = 10, 16
    M = rng.standard_normal((n, n))
    P = sparse.csc_matrix(M.T @ M + 0.1 * np.eye(n))
    q = rng.standard_normal(n)
    A = sparse.csc_matrix(rng.standard_normal((m, n)) * (rng.random((m, n)) < 0.5))
    a = A[0] + 0.3 * A[1]
    A = sparse.vstack([A, a, -a]).tocsc()
    ax = A @ rng.standard_normal(n)
    lo, up = ax - rng.random(m + 2), ax + rng.random(m + 2)
    lo[-2:], up[-2:] = -np.inf, -1.0
    return sparse.triu(P, format="csc"), q, A, lo, up


def _dinf_problem():
    # dual_infeasibility_test.py:25-50 (LP unbounded below)
    P = sparse.csc_matrix((2, 2))
    q = np.array([2.0, -1.0])
    A = sparse.eye(2, format="csc")
    return P, q, A, np.array([0.0, 0.0]), np.array([np.inf, np.inf])


@pytest.mark.parametrize("kind", ["solved", "primal", "dual"])
def test_inaccurate_statuses_match_oracle(kind):
    from oracle import osqp_oracle as oo
    if kind == "solved":
        c = load_case("test_basic_QP")
        prob_data = (c["P"], c["q"], c["A"], c["l"], c["u"])
        st = dict(eps_abs=1e-7, eps_rel=1e-7, check_termination=0, adaptive_rho=0, rho=0.1)
        want, exact = oo.SOLVED_INACCURATE, oo.SOLVED
    elif kind == "primal":
        prob_data = _pinf_problem()
        st = dict(eps_abs=1e-9, eps_rel=1e-9, eps_prim_inf=1e-4, check_termination=0, adaptive_rho=1, adaptive_rho_interval=25)
        want, exact = oo.PRIMAL_INFEASIBLE_INACCURATE, oo.PRIMAL_INFEASIBLE
    else:
        prob_data = _dinf_problem()
        st = dict(eps_abs=1e-9, eps_rel=1e-9, eps_dual_inf=1e-7, check_termination=0, adaptive_rho=0)
        want, exact = oo.DUAL_INFEASIBLE_INACCURATE, oo.DUAL_INFEASIBLE
    hits = _find_budget(lambda k: _oracle(*prob_data, max_iter=k, **st), want)
    assert len(hits) >= 1, "the oracle never passes through the 10x-tolerance window"
    k = hits[len(hits) // 2]      # middle of the longest contiguous window: robust against a one-iteration drift
    res = _engine(*prob_data, max_iter=k, **st).solve()
    ora = _oracle(*prob_data, max_iter=k, **st).solve()
    assert ora.info.status_val == want
    assert res.info.status_val == want, (res.info.status, k, hits)
    assert res.info.iter == k
    if kind == "solved":
        assert _rel(res.x, ora.x) <= 1e-6 and abs(res.info.obj_val - ora.info.obj_val) <= 1e-6 * max(1, abs(ora.info.obj_val))
    elif kind == "primal":
        assert res.info.obj_val == pytest.approx(1e30) and np.all(np.isnan(res.x))
        v, w = res.prim_inf_cert / np.linalg.norm(res.prim_inf_cert), ora.prim_inf_cert / np.linalg.norm(ora.prim_inf_cert)
        assert np.max(np.abs(v - w)) <= 1e-4
    else:
        assert res.info.obj_val == pytest.approx(-1e30) and np.all(np.isnan(res.x))
        v, w = res.dual_inf_cert / np.linalg.norm(res.dual_inf_cert), ora.dual_inf_cert / np.linalg.norm(ora.dual_inf_cert)
        assert np.max(np.abs(v - w)) <= 1e-6
    # with a generous budget the same problem ends in the exact status on both sides
    res2 = _engine(*prob_data, **dict(st, max_iter=20000, check_termination=25)).solve()
    assert res2.info.status_val == exact


# ------------------------------------------------------------------------------------------------ duality gap
@pytest.mark.parametrize("name", ["test_basic_QP", "test_solve", "c1_random_qp"])
def test_dual_objective_and_gap(name):
    c = load_case(name)
    data = (c["P"], c["q"], c["A"], c["l"], c["u"])
    st = dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=100000, check_termination=1, adaptive_rho_interval=50)
    out = {}
    for dg in (0, 1):
        res = _engine(*data, check_dualgap=dg, **st).solve()
        ora = _oracle(*data, check_dualgap=dg, **st).solve()
        assert res.info.status == ora.info.status == "solved"
        # same algorithm: the extra gap test changes the stopping iteration the same way on both sides
        assert abs(res.info.iter - ora.info.iter) <= max(3, 0.2 * ora.info.iter)
        sc = max(1.0, abs(ora.info.obj_val))
        assert abs(res.info.dual_obj_val - ora.info.dual_obj_val) <= 1e-6 * sc
        assert abs(res.info.duality_gap - (res.info.obj_val - res.info.dual_obj_val)) <= 1e-12 * sc
        # strong duality at the optimum
        assert abs(res.info.duality_gap) <= 1e-6 * sc
        # independent evaluation of the Lagrange dual at the returned (x, y)
        Pf = c["P"] + sparse.triu(c["P"], 1).T if sparse.tril(c["P"], -1).nnz == 0 else c["P"]
        y, lo, up = res.y, c["l"], c["u"]
        sup = np.sum(np.where(up < 1e26, up * np.maximum(y, 0), 0.0)) + np.sum(np.where(lo > -1e26, lo * np.minimum(y, 0), 0.0))
        dual = -0.5 * res.x @ (Pf @ res.x) - sup
        assert abs(res.info.dual_obj_val - dual) <= 1e-6 * sc
        out[dg] = res.info.iter
    assert out[1] >= out[0]


def test_dualgap_delays_termination_like_the_oracle():
    """At a loose tolerance the gap test is the last to pass: iteration counts with check_dualgap = 1 equal the oracle's."""
    from b200qp import problems
    P, q, A, l, u = problems.random_qp(400, 600, nnz_per_row=6, seed=21)
    st = dict(eps_abs=1e-4, eps_rel=1e-4, max_iter=20000, check_termination=1, adaptive_rho=0)
    it = {}
    for dg in (0, 1):
        res = _engine(P, q, A, l, u, check_dualgap=dg, **st).solve()
        ora = _oracle(P, q, A, l, u, check_dualgap=dg, **st).solve()
        assert res.info.status == ora.info.status == "solved"
        assert abs(res.info.iter - ora.info.iter) <= 2, (dg, res.info.iter, ora.info.iter)
        it[dg] = res.info.iter
    assert it[1] >= it[0]


# ------------------------------------------------------------------------------------------------ polish
POLISH_CASES = ["test_polish_simple", "test_polish_unconstrained", "test_polish_random"]


@pytest.mark.parametrize("name", POLISH_CASES)
def test_polish_matches_factorisation_polish(name):
    """status_polish == 1 and the polished point within 1e-7 of the oracle's (KKT factorisation + refinement,
    _osqp.py:1693-1828)."""
    c = load_case(name)
    data = (c["P"], c["q"], c["A"], c["l"], c["u"])
    st = dict(c["settings"])
    res = _engine(*data, **st).solve()
    ora = _oracle(*data, linsys="direct", **st).solve()
    assert ora.info.status_polish == 1 and res.info.status_polish == 1
    assert _rel(res.x, ora.x) <= 1e-7
    if c["m"]:
        assert _rel(res.y, ora.y) <= 1e-7
    assert abs(res.info.obj_val - ora.info.obj_val) <= 1e-7 * max(1.0, abs(ora.info.obj_val))
    assert res.info.prim_res <= 1e-8 and res.info.dual_res <= 1e-7


@pytest.mark.parametrize("gen", ["random_tight", "random_loose", "svm", "lasso"])
def test_polish_status_follows_the_reference_algorithm(gen):
    """Where the reference's polish succeeds ours does, to 1e-7; where its active-set guess is wrong (loose tolerances,
    degenerate SVM / Lasso problems) it reports -1 and so do we -- never a 'success' the reference would not report."""
    from b200qp import problems
    if gen == "random_tight":
        data, st = problems.random_qp(300, 450, nnz_per_row=6, seed=5), dict(eps_abs=1e-6, eps_rel=1e-6, max_iter=20000)
    elif gen == "random_loose":
        data, st = problems.random_qp(1000, 1500, nnz_per_row=10, seed=2), dict()
    elif gen == "svm":
        data, st = problems.svm_qp(500, 2000, density=0.02, seed=2, box_rows=1000), dict(adaptive_rho=1)
    else:
        data, st = problems.lasso_qp(300, 600, density=0.05, seed=1), dict()
    res = _engine(*data, polishing=1, **st).solve()
    ora = _oracle(*data, linsys="direct", polishing=1, **st).solve()
    assert res.info.status == ora.info.status == "solved"
    assert res.info.status_polish == ora.info.status_polish, (res.info.status_polish, ora.info.status_polish)
    if ora.info.status_polish == 1:
        assert _rel(res.x, ora.x) <= 1e-7 and _rel(res.y, ora.y) <= 1e-7
    print(f"polish[{gen}]: engine {res.info.status_polish} oracle {ora.info.status_polish} "
          f"polish_time {res.info.polish_time * 1e3:.2f} ms")


# ------------------------------------------------------------------------------------------------ ADVICE r1
@pytest.mark.parametrize("fmt", ["csl", "csr"])
def test_time_limit_is_enforced_without_termination_checks(fmt, monkeypatch):
    """check_termination = 0 and a huge max_iter: the clock alone must stop the solve (every kernel samples it every
    iteration; upstream osqp_solve checks it every iteration)."""
    from b200qp import problems
    if fmt == "csr":
        monkeypatch.setenv("B200QP_FORMAT", "csr")
    P, q, A, l, u = problems.random_qp(20000, 30000, nnz_per_row=8, seed=13)
    prob = _engine(P, q, A, l, u, eps_abs=1e-14, eps_rel=1e-14, max_iter=2_000_000_000, check_termination=0, time_limit=0.2)
    t0 = time.perf_counter()
    res = prob.solve()
    dt = time.perf_counter() - t0
    assert res.info.status_val == prob.constant("OSQP_TIME_LIMIT_REACHED")
    assert dt < 5.0 and 0 < res.info.iter < 2_000_000_000
    assert np.all(np.isfinite(res.x)) and res.info.prim_res > 0


def test_partial_warm_start_keeps_the_other_iterate():
    c = load_case("test_warm_start")
    st = dict(eps_abs=1e-8, eps_rel=1e-8, adaptive_rho=False, polishing=False, check_termination=1)
    prob = _engine(c["P"], c["q"], c["A"], c["l"], c["u"], **st)
    r0 = prob.solve()
    # x only: y keeps the optimal multipliers, so the solve restarts at the optimum
    prob.warm_start(x=r0.x)
    assert prob.solve().info.iter < 10
    # y only after a cold x: not the cold-start count, not immediate
    prob.warm_start(x=np.zeros(c["n"]), y=np.zeros(c["m"]))
    cold = prob.solve().info.iter
    assert cold == r0.info.iter
    prob.warm_start(x=np.zeros(c["n"]))
    prob.warm_start(y=r0.y)
    assert prob.solve().info.iter < cold
    with pytest.raises(Exception):
        prob._solver.warm_start(np.zeros(c["n"] + 1), None)
    with pytest.raises(Exception):
        prob._solver.update_data_vec(q=np.zeros(c["n"] + 3), l=None, u=None)


def test_rejected_bound_update_leaves_q_untouched():
    import osqp
    c = load_case("test_basic_QP")
    prob = _engine(c["P"], c["q"], c["A"], c["l"], c["u"], **dict(c["settings"]))
    base = prob.solve()
    rc = prob._solver.update_data_vec(q=100.0 * np.ones(c["n"]), l=np.ones(c["m"]), u=np.zeros(c["m"]))
    assert rc == osqp.SolverError.OSQP_DATA_VALIDATION_ERROR
    again = prob.solve()      # warm-started re-solve of the SAME problem (q = 100 would move the objective by O(100))
    assert abs(again.info.obj_val - base.info.obj_val) <= 1e-5 * max(1, abs(base.info.obj_val))


def test_non_convex_returns_nan_solution():
    P = sparse.triu([[2.0, 5.0], [5.0, 1.0]], format="csc")
    A = sparse.csc_matrix([[-1.0, 0.0], [0.0, -1.0], [-1.0, 3.0], [2.0, 5.0], [3.0, 4]])
    u = np.array([0.0, 0.0, -15, 100, 80])
    prob = _engine(P, np.array([3.0, 4.0]), A, -np.inf * np.ones(5), u, sigma=5)
    res = prob.solve()
    assert res.info.status_val == prob.constant("OSQP_NON_CVX")
    assert np.all(np.isnan(res.x)) and np.all(np.isnan(res.y)) and np.isnan(res.info.obj_val)


# ------------------------------------------------------------------------------------------------ named sizes
def _first_iterates_parity(P, q, A, l, u, iters, tol, **st):
    """Engine vs oracle after `iters` ADMM iterations with every inner system solved to 1e-9 on both sides."""
    kw = dict(max_iter=iters, check_termination=0, adaptive_rho=0, cg_tol_fraction=1e-9, cg_max_iter=3000, **st)
    t0 = time.perf_counter()
    res = _engine(P, q, A, l, u, **kw).solve()
    t1 = time.perf_counter()
    ora = _oracle(P, q, A, l, u, **dict(kw, rho_eq_factor=10.0)).solve()
    t2 = time.perf_counter()
    assert res.info.iter == ora.info.iter == iters
    ex, ey = _rel(res.x, ora.x), _rel(res.y, ora.y)
    eo = abs(res.info.obj_val - ora.info.obj_val) / max(1.0, abs(ora.info.obj_val))
    ep = abs(res.info.prim_res - ora.info.prim_res) / max(1e-12, ora.info.prim_res)
    ed = abs(res.info.dual_res - ora.info.dual_res) / max(1e-12, ora.info.dual_res)
    print(f"first {iters} iterates: x {ex:.1e} y {ey:.1e} obj {eo:.1e} prim_res {ep:.1e} dual_res {ed:.1e} "
          f"(engine {t1 - t0:.1f} s incl. setup, oracle {t2 - t1:.1f} s)")
    assert ex <= tol and ey <= tol and eo <= tol
    assert ep <= 1e-4 and ed <= 1e-4
    return res, ora


def test_c4_full_size_iterates_vs_oracle():
    """BASELINE config 4 (n = m = 1e6, nnz(A) = 1e7): the first ADMM iterates equal the oracle's to 1e-6."""
    from b200qp import problems
    P, q, A, l, u = problems.random_qp(1_000_000, 1_000_000, nnz_per_row=10, seed=3)
    _first_iterates_parity(P, q, A, l, u, iters=4, tol=1e-6)


def test_c4_full_size_default_solve_vs_oracle():
    """The benchmark solve itself (eps 1e-3, engine defaults): same status, ADMM iteration count within 2 of the oracle's
    engine-spec mode, x and objective within 1e-4 (two inexact-PCG trajectories stopped at a 1e-3 tolerance)."""
    from b200qp import problems
    P, q, A, l, u = problems.random_qp(1_000_000, 1_000_000, nnz_per_row=10, seed=3)
    st = dict(eps_abs=1e-3, eps_rel=1e-3, max_iter=4000)
    prob = _engine(P, q, A, l, u, **st)
    res = prob.solve()
    t0 = time.perf_counter()
    ora = _oracle(P, q, A, l, u, **st).solve(max_wall=600)
    print(f"C4 default solve: engine {res.info.iter} it ({prob.stats()['cg_iters']} PCG) oracle {ora.info.iter} it "
          f"({ora.info.cg_iters} PCG, {time.perf_counter() - t0:.0f} s CPU)")
    assert res.info.status == ora.info.status == "solved"
    assert abs(res.info.iter - ora.info.iter) <= 2
    assert _rel(res.x, ora.x) <= 1e-4
    assert abs(res.info.obj_val - ora.info.obj_val) <= 1e-4 * max(1.0, abs(ora.info.obj_val))


def test_c4_family_tight_vs_oracle():
    """Config-4 family at n = m = 1e5, eps 1e-9: x*, y*, objective within 1e-6 of the oracle (north_star tolerance)."""
    from b200qp import problems
    P, q, A, l, u = problems.random_qp(100_000, 100_000, nnz_per_row=10, seed=3)
    st = dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=20000)
    res = _engine(P, q, A, l, u, **st).solve()
    ora = _oracle(P, q, A, l, u, **st).solve(max_wall=600)
    assert res.info.status == ora.info.status == "solved"
    assert abs(res.info.iter - ora.info.iter) <= max(3, 0.05 * ora.info.iter)
    assert _rel(res.x, ora.x) <= 1e-6 and _rel(res.y, ora.y) <= 1e-6
    assert abs(res.info.obj_val - ora.info.obj_val) <= 1e-6 * max(1.0, abs(ora.info.obj_val))


def test_c2_lasso_full_size_iterates_vs_oracle():
    """BASELINE config 2 (Lasso n = 1e5, m = 2e5, 6.3e7 nnz): first ADMM iterates vs the oracle, exact inner solves."""
    from b200qp import problems
    P, q, A, l, u = problems.lasso_qp(25_000, 50_000, density=0.05, seed=1, box=True)
    _first_iterates_parity(P, q, A, l, u, iters=3, tol=1e-6)


def test_c3_svm_full_size_iterates_vs_oracle():
    """BASELINE config 3 (SVM n = 5e4, m = 1e5): first ADMM iterates vs the oracle, exact inner solves."""
    from b200qp import problems
    P, q, A, l, u = problems.svm_qp(10_000, 40_000, density=0.01, seed=2, box_rows=20_000)
    _first_iterates_parity(P, q, A, l, u, iters=5, tol=1e-6)


# ------------------------------------------------------------------------------------------------ tile ring
def test_second_solver_in_the_process_ring_stays_in_step():
    """Regression for the tile ring (admm_kernel_csl.cuh, consumer loop): an mbarrier wait sees one parity bit, and a
    consumer warp that came for use k of a stage while the fill of use k - 1 was still in flight walked into a stale
    tile -- a hang or an illegal address, seen only on the SECOND solver of a process (bench.py's banded line after
    the random one) once the gathers had become fast.  Config-4 size, random then banded, both kept alive; identical
    cold solves must give identical results and the iteration counts of the single-solver runs."""
    import osqp
    from b200qp import problems
    kw = dict(eps_abs=1e-3, eps_rel=1e-3, max_iter=4000, warm_starting=False, polishing=False, verbose=False)
    keep = []
    for band, iters in ((None, 75), (64, 75), (None, 75)):
        P, q, A, l, u = problems.random_qp(1_000_000, 1_000_000, nnz_per_row=10, seed=3, band=band)
        prob = osqp.OSQP(algebra="cuda")
        prob.setup(P, q, A, l, u, **kw)
        first = prob.solve()
        assert first.info.status == "solved" and first.info.iter == iters, (band, first.info.status, first.info.iter)
        for _ in range(2):
            again = prob.solve()
            assert again.info.status == "solved" and again.info.iter == first.info.iter
            np.testing.assert_array_equal(again.x, first.x)
        keep.append(prob)
