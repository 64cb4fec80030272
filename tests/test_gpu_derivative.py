"""GPU tests of the adjoint derivatives (reference surface src/osqp/interface.py:519-598; cases restated from
src/osqp/tests/derivative_test.py, which needs the reference's builtin algebra and is skipped without it).
osqp.OSQP.adjoint_derivative_compute / _get_mat / _get_vec -> pybind11 -> b200qp_adjoint_derivative_* -> the
persistent ADMM kernel run on the adjoint KKT problem.

Checked against (a) the oracle's active-set adjoint evaluated at the engine's own solution (tight: same formula,
independent dense solve), (b) finite differences of the loss through complete engine solves, at the reference
test's tolerances (rel 5e-3, abs 5e-3; derivative_test.py:15-18)."""
import numpy as np
import numpy.random as npr
import numpy.testing as npt
import pytest
from scipy import sparse
from scipy.optimize import approx_fprime

pytestmark = pytest.mark.gpu

KW = dict(eps_abs=1e-9, eps_rel=1e-9, max_iter=500000, verbose=False)      # derivative_test.py:21-23
REL, ABS, GRAD_EPS = 5e-3, 5e-3, 1e-5                                       # derivative_test.py:14-16


def get_prob(n=10, m=3, seed=1):                                            # derivative_test.py:32-47
    npr.seed(seed)
    L = npr.randn(n, n - 1)
    P = sparse.csc_matrix(L.dot(L.T) + 0.1 * sparse.eye(n))
    x_0 = npr.randn(n)
    s_0 = npr.rand(m)
    A = sparse.csc_matrix(npr.randn(m, n))
    u = A.dot(x_0) + s_0
    l = A.dot(x_0) - s_0
    q = npr.randn(n)
    return P, q, A, l, u, npr.randn(n), npr.randn(m)


def csc_index(M):
    """(rows, cols) in the storage order of M.data.  The reference test uses M.nonzero() for this
    (derivative_test.py:342, 386), which newer scipy returns in row-major order -- not the order of .data."""
    M = sparse.csc_matrix(M)
    return M.indices.copy(), np.repeat(np.arange(M.shape[1]), np.diff(M.indptr))


def solve(P, q, A, l, u):
    import osqp
    prob = osqp.OSQP(algebra="cuda")
    prob.setup(P, q, A, l, u, **KW)
    res = prob.solve()
    assert res.info.status_val == osqp.SolverStatus.OSQP_SOLVED, res.info.status
    return prob, res


def get_grads(P, q, A, l, u, true_x, true_y=None):                          # derivative_test.py:49-75
    prob, res = solve(P, q, A, l, u)
    if true_y is None:
        prob.adjoint_derivative_compute(dx=res.x - true_x)
    else:
        prob.adjoint_derivative_compute(dx=res.x - true_x, dy=res.y - true_y)
    dP, dA = prob.adjoint_derivative_get_mat()
    dq, dl, du = prob.adjoint_derivative_get_vec()
    return (dP, dq, dA, dl, du), res


def loss(P, q, A, l, u, true_x):
    _, res = solve(P, q, A, l, u)
    return 0.5 * np.sum(np.square(res.x - true_x))


@pytest.mark.parametrize("n,m,with_y", [(10, 3, False), (5, 5, True), (30, 30, False), (10, 20, True)])
def test_adjoint_matches_oracle_adjoint(n, m, with_y):
    from oracle.osqp_oracle import adjoint_derivative
    P, q, A, l, u, tx, ty = get_prob(n, m)
    (dP, dq, dA, dl, du), res = get_grads(P, q, A, l, u, tx, ty if with_y else None)
    oP, oq, oA, ol, ou = adjoint_derivative(P, q, A, l, u, res.x, res.y, res.x - tx, (res.y - ty) if with_y else None)
    scale = max(1.0, np.abs(oq).max())
    npt.assert_allclose(dq, oq, rtol=0, atol=1e-5 * scale)
    npt.assert_allclose(dl, ol, rtol=0, atol=1e-5 * max(1.0, np.abs(ol).max()))
    npt.assert_allclose(du, ou, rtol=0, atol=1e-5 * max(1.0, np.abs(ou).max()))
    npt.assert_allclose(np.triu(dP), np.triu(oP), rtol=0, atol=1e-5 * max(1.0, np.abs(oP).max()))
    npt.assert_allclose(dA, oA, rtol=0, atol=1e-5 * max(1.0, np.abs(oA).max()))


def test_dl_dq():                                                            # derivative_test.py:298-335
    P, q, A, l, u, tx, _ = get_prob(5, 5)
    (_, dq, _, _, _), _ = get_grads(P, q, A, l, u, tx)
    dq_fd = approx_fprime(q, lambda v: loss(P, v, A, l, u, tx), GRAD_EPS)
    npt.assert_allclose(dq, dq_fd, rtol=REL, atol=ABS)


def test_dl_dP():                                                            # derivative_test.py:337-379
    P, q, A, l, u, tx, _ = get_prob(3, 3)
    P_idx = csc_index(P)
    (dP, _, _, _, _), _ = get_grads(P, q, A, l, u, tx)
    f = lambda pv: loss(sparse.csc_matrix((pv, P_idx), shape=P.shape), q, A, l, u, tx)
    dP_fd = sparse.csc_matrix((approx_fprime(P.data, f, GRAD_EPS), P_idx), shape=P.shape)
    dP_fd = (dP_fd + dP_fd.T) / 2
    npt.assert_allclose(np.triu(dP), np.triu(dP_fd.todense()), rtol=REL, atol=ABS)


def test_dl_dA():                                                            # derivative_test.py:381-422
    P, q, A, l, u, tx, _ = get_prob(3, 3)
    A_idx = csc_index(A)
    (_, _, dA, _, _), _ = get_grads(P, q, A, l, u, tx)
    f = lambda av: loss(P, q, sparse.csc_matrix((av, A_idx), shape=A.shape), l, u, tx)
    dA_fd = sparse.csc_matrix((approx_fprime(A.data, f, GRAD_EPS), A_idx), shape=A.shape)
    npt.assert_allclose(dA, dA_fd.todense(), rtol=REL, atol=ABS)


def test_dl_dl_and_du():                                                     # derivative_test.py:424-496
    P, q, A, l, u, tx, _ = get_prob(10, 20)
    (_, _, _, dl, du), _ = get_grads(P, q, A, l, u, tx)
    dl_fd = approx_fprime(l, lambda v: loss(P, q, A, v, u, tx), GRAD_EPS)
    du_fd = approx_fprime(u, lambda v: loss(P, q, A, l, v, tx), GRAD_EPS)
    npt.assert_allclose(dl_fd, dl, rtol=REL, atol=ABS)
    npt.assert_allclose(du_fd, du, rtol=REL, atol=ABS)


def test_dl_dA_eq():                                                         # derivative_test.py:498-545
    from oracle.osqp_oracle import adjoint_derivative
    P, q, A, l, u, tx, _ = get_prob(30, 20)
    u[:10] = l[:10]
    (dP, dq, dA, dl, du), res = get_grads(P, q, A, l, u, tx)
    _, oq, oA, _, _ = adjoint_derivative(P, q, A, l, u, res.x, res.y, res.x - tx)
    npt.assert_allclose(dA, oA, rtol=REL, atol=ABS)
    npt.assert_allclose(dq, oq, rtol=REL, atol=ABS)


def test_derivative_requires_a_solved_problem_and_keeps_the_solution():
    import osqp
    P, q, A, l, u, tx, _ = get_prob(10, 3)
    prob = osqp.OSQP(algebra="cuda")
    prob.setup(P, q, A, l, u, **KW)
    assert prob.has_capability("OSQP_CAPABILITY_DERIVATIVES")
    with pytest.raises(ValueError):
        prob.adjoint_derivative_compute(dx=np.ones(10))                      # interface.py:526-531
    r1 = prob.solve()
    prob.adjoint_derivative_compute(dx=r1.x - tx)
    dq1, _, _ = prob.adjoint_derivative_get_vec()
    dPs, dAs = prob.adjoint_derivative_get_mat(as_dense=False, dP_as_triu=False)
    assert sparse.issparse(dPs) and abs(dPs - dPs.T).max() < 1e-14 and dAs.shape == A.shape
    # the solver's own iterates are untouched: a warm-started solve stops at its first check
    r2 = prob.solve()
    assert r2.info.iter <= 25 and np.allclose(r2.x, r1.x, atol=1e-8)
    prob.update(q=q + 1.0)
    with pytest.raises(ValueError):
        prob.adjoint_derivative_get_vec()                                    # results forgotten, interface.py:364-368
