"""GPU parity tests of the batched path (BASELINE.json configs[4]): osqp.OSQPBatch -> osqp_cuda.OSQPBatchSolver
(pybind11) -> b200qp_batch_* C ABI -> admm_batch_kernel (one CTA per QP, vectors in shared memory).

Element b of a batch must be the QP a reference user gets from setup(P, q[0], A, l[0], u[0]), update(q=q[b],
l=l[b], u=u[b]), solve() (src/osqp/nn/torch.py:128-164 re-uses solvers exactly like that).  Checked against
  * the CPU oracle run live through that same call sequence (direct mode, the pinned one): x*, y*, obj within
    1e-6 relative at eps 1e-9 (BASELINE.json north_star);
  * the single-QP engine (osqp.OSQP) through that call sequence: same status and iteration count (PCG mode),
    and the oracle's direct mode with the engine's rho ratio: same iteration count (on-chip K^-1 mode);
  * size-independent properties at the full 8192 x (n=120, m=240): every element solved, KKT residuals of the
    returned (x, y) below the termination tolerance, bitwise determinism between two runs.
"""
import numpy as np
import pytest
from scipy import sparse

pytestmark = pytest.mark.gpu

REL = 1e-6


@pytest.fixture(params=["dense", "pcg"])
def mode(request, monkeypatch):
    """Both inner solvers of the batch kernel: K^-1 in shared memory (default for small QPs) and Jacobi-PCG."""
    monkeypatch.setenv("B200QP_BATCH_DENSE", "1" if request.param == "dense" else "0")
    return request.param


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1.0, float(np.max(np.abs(b))))) if len(b) else 0.0


def _oracle_element(P, q0, A, l0, u0, q, l, u, **kw):
    from oracle.osqp_oracle import OracleSolver
    s = OracleSolver(P, q0, A, l0, u0, linsys="direct", **kw)
    s.update(q=q, l=l, u=u)
    return s.solve()


def test_mpc_batch_matches_oracle_tight(mode):
    import osqp
    from b200qp import problems
    B = 12
    P, q, A, l, u = problems.mpc_batch(B)
    kw = dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=20000)
    prob = osqp.OSQPBatch()
    prob.setup(P, q, A, l, u, verbose=False, cg_tol_fraction=1e-3, **kw)
    res = prob.solve()
    assert prob.stats()["dense"] == (1 if mode == "dense" else 0)
    assert res.x.shape == (B, 120) and res.y.shape == (B, 240)
    for b in range(B):
        ora = _oracle_element(P, q, A, l[0], u[0], q, l[b], u[b], **kw)
        assert res.info.status[b] == "solved" and ora.info.status == "solved"
        assert _rel(res.x[b], ora.x) < REL, (b, _rel(res.x[b], ora.x))
        assert _rel(res.y[b], ora.y) < 10 * REL, (b, _rel(res.y[b], ora.y))
        assert abs(res.info.obj_val[b] - ora.info.obj_val) <= REL * max(1.0, abs(ora.info.obj_val))


def test_mpc_batch_dense_matches_direct_oracle_iterations():
    """On-chip K^-1 mode is an exact inner solve like the reference's factorisation: same iteration counts and
    rho updates as the oracle's direct mode (run with the engine's equality-row rho ratio)."""
    import osqp
    from b200qp import problems
    B = 16
    P, q, A, l, u = problems.mpc_batch(B)
    prob = osqp.OSQPBatch()
    prob.setup(P, q, A, l, u, verbose=False, warm_starting=False)
    res = prob.solve()
    st = prob.stats()
    assert st["dense"] == 1 and st["block"] in (512, 1024) and st["cg_iters"] == 0 and st["kernel_ms"] > 0
    for b in range(B):
        ora = _oracle_element(P, q, A, l[0], u[0], q, l[b], u[b], rho_eq_factor=10.0)
        assert res.info.status[b] == ora.info.status == "solved"
        assert res.info.iter[b] == ora.info.iter, (b, res.info.iter[b], ora.info.iter)
        assert res.info.rho_updates[b] == ora.info.rho_updates
        assert _rel(res.x[b], ora.x) < 1e-7 and _rel(res.y[b], ora.y) < 1e-7
        assert abs(res.info.obj_val[b] - ora.info.obj_val) < 1e-8 * max(1.0, abs(ora.info.obj_val))


def test_mpc_batch_matches_single_engine(monkeypatch):
    """PCG mode: same iteration counts and results as looping the single-QP engine the way the reference's
    batched caller does: one setup, update(l, u), cold-start solve."""
    import osqp
    from b200qp import problems
    monkeypatch.setenv("B200QP_BATCH_DENSE", "0")
    B = 24
    P, q, A, l, u = problems.mpc_batch(B)
    prob = osqp.OSQPBatch()
    prob.setup(P, q, A, l, u, verbose=False, warm_starting=False)
    res = prob.solve()
    one = osqp.OSQP()
    one.setup(P, q, A, l[0], u[0], verbose=False, warm_starting=False)
    for b in range(B):
        one.update(l=l[b], u=u[b])
        one.update_settings(rho=0.1)          # the batch starts every element from settings.rho
        r1 = one.solve()
        assert res.info.status[b] == r1.info.status == "solved"
        assert res.info.iter[b] == r1.info.iter, (b, res.info.iter[b], r1.info.iter)
        assert res.info.rho_updates[b] == r1.info.rho_updates
        assert _rel(res.x[b], r1.x) < 1e-9 and _rel(res.y[b], r1.y) < 1e-9
        assert abs(res.info.obj_val[b] - r1.info.obj_val) < 1e-9 * max(1.0, abs(r1.info.obj_val))
    st = prob.stats()
    assert st["admm_iters"] == int(res.info.iter.sum()) and st["cg_iters"] == int(res.info.cg_iter.sum())
    assert st["kernel_ms"] > 0 and st["block"] == 256


def test_random_batch_with_q_and_onesided_bounds(mode):
    import osqp
    from b200qp import problems
    B, n, m = 9, 60, 90
    P, q0, A, l0, u0 = problems.random_qp(n, m, nnz_per_row=4, seed=11, eq_frac=0.2, onesided_frac=0.2)
    rng = np.random.default_rng(5)
    q = q0 + 0.3 * rng.standard_normal((B, n))
    shift = 0.2 * rng.standard_normal((B, m))
    l, u = l0 + shift, u0 + shift               # +-1e30 entries stay "infinite" (|.| > 1e26)
    q[0], l[0], u[0] = q0, l0, u0
    kw = dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=50000)
    prob = osqp.OSQPBatch()
    prob.setup(P, q, A, l, u, verbose=False, cg_tol_fraction=1e-3, **kw)
    res = prob.solve()
    for b in range(B):
        ora = _oracle_element(P, q[0], A, l[0], u[0], q[b], np.maximum(l[b], -1e30), np.minimum(u[b], 1e30), **kw)
        assert res.info.status[b] == "solved" and ora.info.status == "solved"
        assert _rel(res.x[b], ora.x) < REL and _rel(res.y[b], ora.y) < 10 * REL
        assert abs(res.info.obj_val[b] - ora.info.obj_val) <= REL * max(1.0, abs(ora.info.obj_val))


def test_update_and_warm_start(mode):
    import osqp
    from b200qp import problems
    B = 8
    P, q, A, l, u = problems.mpc_batch(B)
    prob = osqp.OSQPBatch()
    prob.setup(P, q, A, l, u, verbose=False)
    r1 = prob.solve()
    assert all(s == "solved" for s in r1.info.status)
    # warm_starting (default on): solving again from the solution needs no more iterations than before
    r2 = prob.solve()
    assert np.all(r2.info.iter <= r1.info.iter) and all(s == "solved" for s in r2.info.status)
    np.testing.assert_allclose(r2.x, r1.x, atol=1e-2)
    # update to another set of initial states == a fresh batch on those states (cold start both)
    P2, q2, A2, l2, u2 = problems.mpc_batch(B, seed0=100)
    prob.update_settings(warm_starting=False)
    prob.update_settings(rho=0.1)
    prob.update(l=l2, u=u2)
    r3 = prob.solve()
    fresh = osqp.OSQPBatch()
    fresh.setup(P, q, A, np.vstack([l[:1], l2[1:]]), np.vstack([u[:1], u2[1:]]), verbose=False, warm_starting=False)
    r4 = fresh.solve()
    np.testing.assert_array_equal(r3.info.iter[1:], r4.info.iter[1:])
    np.testing.assert_allclose(r3.x[1:], r4.x[1:], rtol=0, atol=1e-12)
    # explicit warm start from the known solution: one check interval
    prob.warm_start(x=r3.x, y=r3.y)
    r5 = prob.solve()
    assert np.all(r5.info.iter <= r3.info.iter)
    with pytest.raises(osqp.OSQPException) as ei:
        prob.update(l=u2 + 1.0, u=u2)
    assert ei.value == osqp.SolverError.OSQP_DATA_VALIDATION_ERROR


def test_infeasible_element_is_isolated(mode):
    """One primal-infeasible and one dual-infeasible element must not disturb their neighbours."""
    import osqp
    n = 2
    P = sparse.diags([1.0, 0.0], format="csc")              # x2 has no curvature
    A = sparse.csc_matrix(np.array([[1.0, 0.0], [1.0, 0.0], [0.0, 1.0]]))
    q = np.array([[1.0, 1.0], [1.0, 1.0], [1.0, -1.0], [1.0, 1.0]])
    l = np.array([[0.0, -1.0, 0.0], [2.0, -5.0, 0.0], [0.0, -1.0, 0.0], [0.5, -1.0, 0.5]])
    u = np.array([[1.0, 1.0, 1.0], [5.0, 1.0, 1.0], [1.0, 1.0, 1e30], [1.0, 1.0, 1.0]])
    prob = osqp.OSQPBatch()
    prob.setup(P, q, A, l, u, verbose=False, eps_abs=1e-6, eps_rel=1e-6)
    res = prob.solve()
    assert res.info.status[0] == "solved" and res.info.status[3] == "solved"
    assert res.info.status[1] == "primal infeasible"         # x1 >= 2 and x1 <= 1
    assert res.info.status[2] == "dual infeasible"           # min -x2, x2 unbounded above
    np.testing.assert_allclose(res.x[0], [0.0, 0.0], atol=1e-5)
    np.testing.assert_allclose(res.x[3], [0.5, 0.5], atol=1e-5)
    assert np.all(np.isnan(res.x[1])) and np.all(np.isnan(res.y[1])) and res.info.obj_val[1] == 1e30
    assert np.all(np.isnan(res.x[2])) and res.info.obj_val[2] == -1e30
    pc = res.prim_inf_cert[1]
    assert np.all(np.isfinite(pc)) and pc[0] * pc[1] < 0      # the two conflicting rows, opposite signs
    dc = res.dual_inf_cert[2]
    assert np.all(np.isfinite(dc)) and dc[1] > 0 and abs(dc[0]) < 1e-3 * abs(dc[1])
    assert np.all(np.isnan(res.prim_inf_cert[0])) and np.all(np.isnan(res.dual_inf_cert[0]))


def test_full_config5_properties(mode):
    """8192 x (n = 120, m = 240): size-independent checks."""
    import osqp
    from b200qp import problems
    B = 8192
    P, q, A, l, u = problems.mpc_batch(B)
    eps = 1e-4
    prob = osqp.OSQPBatch()
    prob.setup(P, q, A, l, u, verbose=False, eps_abs=eps, eps_rel=eps, warm_starting=False)
    res = prob.solve()
    assert np.all(res.info.status_val == 1), np.unique(res.info.status_val, return_counts=True)
    assert np.all(res.info.iter % 25 == 0) and res.info.iter.max() <= 4000
    Pf = sparse.csr_matrix(P)
    Ac = sparse.csr_matrix(A)
    ax = res.x @ Ac.T                                   # (B, m)
    prim = np.maximum(np.maximum(l - ax, ax - u), 0).max(axis=1)
    dual = np.abs(res.x @ Pf.T + q + res.y @ Ac).max(axis=1)
    scale_p = np.maximum(np.abs(ax).max(axis=1), 1.0)
    scale_d = np.maximum(np.maximum(np.abs(res.x @ Pf.T).max(axis=1), np.abs(res.y @ Ac).max(axis=1)), 1.0)
    assert np.all(prim <= 2 * eps * (1 + scale_p)), prim.max()
    assert np.all(dual <= 2 * eps * (1 + scale_d)), dual.max()
    obj = 0.5 * np.einsum("bi,bi->b", res.x @ Pf.T, res.x) + res.x @ q
    np.testing.assert_allclose(res.info.obj_val, obj, rtol=1e-9, atol=1e-9)
    # a sample against the oracle at the same tolerance (different inner solver: compare to the tolerance)
    for b in (0, 1, 4097, 8191):
        ora = _oracle_element(P, q, A, l[0], u[0], q, l[b], u[b], eps_abs=eps, eps_rel=eps)
        assert ora.info.status == "solved"
        assert _rel(res.x[b], ora.x) < 5e-3 and abs(res.info.obj_val[b] - ora.info.obj_val) < 5e-3 * max(1.0, abs(ora.info.obj_val))
    # determinism: a second cold-start solve is bitwise identical
    prob.update_settings(rho=0.1)
    res2 = prob.solve()
    np.testing.assert_array_equal(res.info.iter, res2.info.iter)
    assert np.array_equal(res.x, res2.x) and np.array_equal(res.y, res2.y)


def test_batch_too_large_for_shared_memory_fails_loudly():
    import osqp
    from b200qp import problems
    P, q, A, l, u = problems.random_qp(3000, 3000, nnz_per_row=3, seed=1)
    prob = osqp.OSQPBatch()
    with pytest.raises(osqp.OSQPException) as ei:
        prob.setup(P, q, A, l, u, batch=2, verbose=False)
    assert ei.value == osqp.SolverError.OSQP_MEM_ALLOC_ERROR


def test_mid_size_batch_falls_back_to_pcg():
    """n = 400 does not fit K^-1 in shared memory: the batch uses the PCG kernel and still matches the oracle."""
    import osqp
    from b200qp import problems
    B, n, m = 5, 400, 500
    P, q0, A, l0, u0 = problems.random_qp(n, m, nnz_per_row=5, seed=21, eq_frac=0.1)
    rng = np.random.default_rng(8)
    q = q0 + 0.1 * rng.standard_normal((B, n))
    kw = dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=50000)
    prob = osqp.OSQPBatch()
    prob.setup(P, q, A, l0, u0, verbose=False, cg_tol_fraction=1e-3, **kw)
    res = prob.solve()
    assert prob.stats()["dense"] == 0 and prob.stats()["block"] == 256
    for b in range(B):
        ora = _oracle_element(P, q[0], A, l0, u0, q[b], l0, u0, **kw)
        assert res.info.status[b] == "solved" and ora.info.status == "solved"
        assert _rel(res.x[b], ora.x) < REL and _rel(res.y[b], ora.y) < 10 * REL


def _perturbed_matrices(P, A, B, rng):
    """Per-element values on the shared pattern: P_b = P + diag(nonnegative), A_b = A .* (1 + small)."""
    Pt = sparse.triu(P, format="csc")
    Pt.sort_indices()
    Ac = sparse.csc_matrix(A)
    Ac.sort_indices()
    Px = np.tile(Pt.data, (B, 1))
    rows = Pt.indices
    cols = np.repeat(np.arange(Pt.shape[1]), np.diff(Pt.indptr))
    diag = rows == cols
    Px[:, diag] += 0.5 * rng.random((B, int(diag.sum())))
    Ax = np.tile(Ac.data, (B, 1)) * (1.0 + 0.02 * rng.standard_normal((B, Ac.nnz)))
    Px[0], Ax[0] = Pt.data, Ac.data
    return Pt, Ac, Px, Ax


def test_per_element_matrix_values(mode):
    """Element b = setup(element 0) + update(Px=Px[b], Ax=Ax[b], q=q[b]) + solve(): the reference's cached-solver
    path for per-element P_val / A_val (src/osqp/nn/torch.py:136-140)."""
    import osqp
    from b200qp import problems
    from oracle.osqp_oracle import OracleSolver
    B, n, m = 7, 40, 60
    P, q0, A, l, u = problems.random_qp(n, m, nnz_per_row=4, seed=31, eq_frac=0.0)
    rng = np.random.default_rng(3)
    Pt, Ac, Px, Ax = _perturbed_matrices(P, A, B, rng)
    q = q0 + 0.2 * rng.standard_normal((B, n))
    q[0] = q0
    l, u = l - 0.5, u + 0.5                       # keep every perturbed element feasible
    kw = dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=50000)
    prob = osqp.OSQPBatch()
    prob.setup(P, q, A, l, u, Px=Px, Ax=Ax, verbose=False, cg_tol_fraction=1e-3, **kw)
    res = prob.solve()
    for b in range(B):
        s = OracleSolver(P, q0, A, l, u, linsys="direct", **kw)
        s.update(q=q[b], Px=Px[b], Ax=Ax[b], P_pattern=Pt, A_pattern=Ac)
        ora = s.solve()
        assert res.info.status[b] == ora.info.status == "solved", (b, res.info.status[b], ora.info.status)
        assert _rel(res.x[b], ora.x) < REL and _rel(res.y[b], ora.y) < 10 * REL, (b, _rel(res.x[b], ora.x), _rel(res.y[b], ora.y))
        assert abs(res.info.obj_val[b] - ora.info.obj_val) <= REL * max(1.0, abs(ora.info.obj_val))
    # updating only A afterwards keeps the per-element P; back to element 0's A everywhere
    prob.update(Ax=Ac.data)
    prob.update_settings(rho=0.1)
    prob.update_settings(warm_starting=False)
    res2 = prob.solve()
    s = OracleSolver(P, q0, A, l, u, linsys="direct", **kw)
    s.update(q=q[3], Px=Px[3], P_pattern=Pt)
    ora = s.solve()
    assert _rel(res2.x[3], ora.x) < REL and abs(res2.info.obj_val[3] - ora.info.obj_val) <= REL * max(1.0, abs(ora.info.obj_val))
    with pytest.raises(osqp.OSQPException):
        prob.update(Ax=np.zeros((B, Ac.nnz + 1)))


def test_per_element_values_equal_to_shared_give_the_same_solves():
    """Same numbers through the per-element path (values re-scaled as E a D in one product instead of the ten Ruiz
    passes of setup: equal up to rounding)."""
    import osqp
    from b200qp import problems
    B = 6
    P, q, A, l, u = problems.mpc_batch(B)
    shared = osqp.OSQPBatch()
    shared.setup(P, q, A, l, u, verbose=False, warm_starting=False)
    r1 = shared.solve()
    Pt = sparse.triu(P, format="csc"); Pt.sort_indices()
    Ac = sparse.csc_matrix(A); Ac.sort_indices()
    per = osqp.OSQPBatch()
    per.setup(P, q, A, l, u, Px=np.tile(Pt.data, (B, 1)), Ax=np.tile(Ac.data, (B, 1)), verbose=False, warm_starting=False)
    r2 = per.solve()
    np.testing.assert_array_equal(r1.info.iter, r2.info.iter)
    np.testing.assert_allclose(r2.x, r1.x, rtol=0, atol=1e-9)
    np.testing.assert_allclose(r2.y, r1.y, rtol=0, atol=1e-9)
    assert per.stats()["dense"] == 1


def test_unconstrained_and_tiny_batches(mode):
    """m = 0 (reference unconstrained_test.py) and n = 1..3: x* = -P^-1 q for every element."""
    import osqp
    rng = np.random.default_rng(12)
    for n in (1, 3, 30):
        M = rng.standard_normal((n, n))
        P = sparse.csc_matrix(M @ M.T + np.eye(n))
        B = 5
        q = rng.standard_normal((B, n))
        prob = osqp.OSQPBatch()
        prob.setup(P, q, None, None, None, batch=B, verbose=False, eps_abs=1e-9, eps_rel=1e-9)
        res = prob.solve()
        assert all(s == "solved" for s in res.info.status)
        want = -np.linalg.solve(P.toarray(), q.T).T
        np.testing.assert_allclose(res.x, want, rtol=0, atol=1e-6 * max(1.0, np.abs(want).max()))
        assert res.y.shape == (B, 0)


def test_non_convex_element_is_reported(mode):
    """Indefinite P: the reference's indirect path must end NON_CVX or MAX_ITER_REACHED (non_convex_test.py:22-56);
    a convex neighbour in the same batch is not possible (P is shared), so the whole batch reports it."""
    import osqp
    P = sparse.diags([1.0, -1.0], format="csc")
    A = sparse.eye(2, format="csc")
    q = np.array([[0.1, 0.2], [-0.3, 0.1]])
    prob = osqp.OSQPBatch()
    prob.setup(P, q, A, -np.ones(2), np.ones(2), batch=2, verbose=False, max_iter=200)
    res = prob.solve()
    assert all(int(s) in (7, 9) for s in res.info.status_val), res.info.status
    if np.all(res.info.status_val == 9):
        assert np.all(np.isnan(res.x)) and np.all(np.isnan(res.info.obj_val))


def test_batch_with_long_rows_uses_the_16_lane_kernels(mode):
    """Rows with > 12 entries on average select 16 lanes per row (both matrices): same answers as the oracle."""
    import osqp
    from b200qp import problems
    B, n, m = 4, 50, 60
    P, q0, A, l, u = problems.random_qp(n, m, nnz_per_row=30, seed=41, eq_frac=0.1, p_pairs=8)
    assert A.nnz / m > 12 and (sparse.csc_matrix(P).nnz + A.nnz) / n / 2 > 12
    rng = np.random.default_rng(9)
    q = q0 + 0.1 * rng.standard_normal((B, n))
    q[0] = q0
    kw = dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=50000)
    prob = osqp.OSQPBatch()
    prob.setup(P, q, A, l, u, verbose=False, cg_tol_fraction=1e-3, **kw)
    res = prob.solve()
    for b in range(B):
        ora = _oracle_element(P, q[0], A, l, u, q[b], l, u, **kw)
        assert res.info.status[b] == "solved" and ora.info.status == "solved"
        assert _rel(res.x[b], ora.x) < REL and _rel(res.y[b], ora.y) < 10 * REL
        assert abs(res.info.obj_val[b] - ora.info.obj_val) <= REL * max(1.0, abs(ora.info.obj_val))
