"""GPU tests of the multi-GPU path (csrc/admm_kernel_dist.cuh: row slabs + variable slices, all-gather by stores).

On a one-GPU box the ranks share device 0 (each rank is its own cooperative kernel on its own stream;
the cross-rank protocol -- flags, payloads, pushes through peer pointers -- is exactly the multi-device one).  With >= 2 visible GPUs the same tests also run one rank per device."""
import numpy as np
import pytest

from case_runner import load_case, run_case_engine, run_case_oracle

pytestmark = pytest.mark.gpu


def _devices(world):
    import osqp_cuda
    nd = osqp_cuda.device_count()
    lists = [[0] * world]
    if nd >= world and world > 1:
        lists.append(list(range(world)))
    return lists


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1.0, float(np.max(np.abs(b))))) if len(b) else 0.0


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("name", ["test_basic_QP", "test_solve", "polish_random_tight", "test_warm_start"])
def test_group_matches_single_gpu_and_oracle(name, world):
    from b200qp.multigpu import LocalGroup
    case = load_case(name)
    st = dict(case["settings"], verbose=False, max_iter=40000, polishing=False)
    single, _ = run_case_engine(case, max_iter=40000, polishing=False)
    ora = run_case_oracle(case, "direct", max_iter=40000, polishing=0)
    for devs in _devices(world):
        grp = LocalGroup(case["P"], case["q"], case["A"], case["l"], case["u"], devs, **st)
        results = grp.solve()
        r0 = results[0]
        for r in results:                      # every rank reports the same thing, bit for bit
            assert r.info.status == "solved" and r.info.iter == r0.info.iter
            np.testing.assert_array_equal(r.x, r0.x)
            assert r.info.obj_val == r0.info.obj_val
        assert abs(r0.info.iter - single.info.iter) <= max(3, 0.05 * single.info.iter)
        for ref in (single, ora):
            assert _rel(r0.x, ref.x) <= 1e-6 and _rel(r0.y, ref.y) <= 1e-5
            assert abs(r0.info.obj_val - ref.info.obj_val) <= 1e-6 * max(1.0, abs(ref.info.obj_val))
        # second solve on the same group: warm-started from the solution, flags/epochs carry over
        again = grp.solve()[0]
        assert again.info.status == "solved" and again.info.iter <= max(25, r0.info.iter // 4)


@pytest.mark.parametrize("world", [2, 4])
def test_group_midsize_random_qp(world):
    """Config-4 family at n = m = 2e4 (full grids on every rank, adaptive rho, default cadence)."""
    import osqp
    from b200qp import problems
    from b200qp.multigpu import LocalGroup
    P, q, A, l, u = problems.random_qp(20000, 20000, nnz_per_row=10, seed=3)
    kw = dict(verbose=False, eps_abs=1e-6, eps_rel=1e-6, max_iter=4000)
    prob = osqp.OSQP()
    prob.setup(P, q, A, l, u, **kw)
    single = prob.solve()
    assert single.info.status == "solved"
    for devs in _devices(world):
        grp = LocalGroup(P, q, A, l, u, devs, **kw)
        r = grp.solve()[0]
        assert r.info.status == "solved"
        assert abs(r.info.iter - single.info.iter) <= max(25, 0.1 * single.info.iter)
        assert _rel(r.x, single.x) <= 1e-4 and _rel(r.y, single.y) <= 1e-3
        rep = problems.kkt_report(P, q, A, l, u, r.x, r.y)
        assert rep["prim"] < 1e-4 and rep["dual"] < 1e-4


def test_group_world8_virtual_ranks():
    """The maximum world size (MAXW = 8) on one device: 8 slabs, 8 slices, 7 peers per push."""
    import osqp
    from b200qp import problems
    from b200qp.multigpu import LocalGroup
    P, q, A, l, u = problems.random_qp(3000, 4000, nnz_per_row=6, seed=5, eq_frac=0.1, onesided_frac=0.1)
    kw = dict(verbose=False, eps_abs=1e-7, eps_rel=1e-7, max_iter=20000)
    prob = osqp.OSQP()
    prob.setup(P, q, A, l, u, **kw)
    single = prob.solve()
    assert single.info.status == "solved"
    grp = LocalGroup(P, q, A, l, u, [0] * 8, **kw)
    results = grp.solve()
    r0 = results[0]
    for r in results:
        assert r.info.status == "solved" and r.info.iter == r0.info.iter
        np.testing.assert_array_equal(r.x, r0.x)
    assert abs(r0.info.iter - single.info.iter) <= max(25, 0.1 * single.info.iter)
    assert _rel(r0.x, single.x) <= 1e-5 and _rel(r0.y, single.y) <= 1e-4
    rep = problems.kkt_report(P, q, A, l, u, r0.x, r0.y)
    assert rep["prim"] < 1e-5 and rep["dual"] < 1e-5


def test_group_infeasible_and_unconstrained():
    from b200qp.multigpu import LocalGroup
    case = load_case("test_primal_infeasibility")
    st = dict(case["settings"], verbose=False)
    r = LocalGroup(case["P"], case["q"], case["A"], case["l"], case["u"], [0, 0], **st).solve()[0]
    assert r.info.status == "primal infeasible"
    v = r.prim_inf_cert / np.linalg.norm(r.prim_inf_cert)
    np.testing.assert_allclose(v, case["expect"]["ref_normalized_prim_inf_cert_correct"], atol=1e-4)
    case = load_case("test_dual_infeasible_qp")
    st = dict(case["settings"], verbose=False)
    r = LocalGroup(case["P"], case["q"], case["A"], case["l"], case["u"], [0, 0], **st).solve()[0]
    assert r.info.status == "dual infeasible"
    case = load_case("test_unconstrained_problem")      # m = 0: every slab is empty
    st = dict(case["settings"], verbose=False)
    r = LocalGroup(case["P"], case["q"], case["A"], case["l"], case["u"], [0, 0], **st).solve()[0]
    assert r.info.status == "solved"
    np.testing.assert_allclose(r.x, case["expect"]["ref_x_val"], rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("fmt,kind", [("tiles", 3), ("csr", 2)])
def test_group_kernel_selection_and_csr_fallback(fmt, kind, monkeypatch):
    """Both multi-GPU kernels behind the same driver: the tile kernel (admm_kernel_csl<true>, default) and the CSR
    fallback (admm_kernel_dist; B200QP_DIST_FORMAT=csr, or matrices the tiles cannot hold) -- same solution, the
    kernel in use reported by stats().  With the tile kernel every rank holds all of y at exit (row range [0, m))."""
    import osqp
    from b200qp import problems
    from b200qp.multigpu import LocalGroup
    if fmt == "csr":
        monkeypatch.setenv("B200QP_DIST_FORMAT", "csr")
    P, q, A, l, u = problems.random_qp(6000, 8000, nnz_per_row=8, seed=11, eq_frac=0.1)
    kw = dict(verbose=False, eps_abs=1e-7, eps_rel=1e-7, max_iter=20000)
    prob = osqp.OSQP()
    prob.setup(P, q, A, l, u, **kw)
    single = prob.solve()
    assert single.info.status == "solved"
    grp = LocalGroup(P, q, A, l, u, [0, 0, 0], **kw)
    results = grp.solve()
    assert all(st["kernel_kind"] == kind for st in grp.stats())
    ranges = [s.dist_row_range() for s in grp.solvers]
    if kind == 3:
        assert all(tuple(rg) == (0, A.shape[0]) for rg in ranges)
    else:
        assert ranges[0][0] == 0 and ranges[-1][1] == A.shape[0] and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    r0 = results[0]
    for r in results:
        assert r.info.status == "solved" and r.info.iter == r0.info.iter
        np.testing.assert_array_equal(r.x, r0.x)
        np.testing.assert_array_equal(r.y, r0.y)
    assert abs(r0.info.iter - single.info.iter) <= max(25, 0.1 * single.info.iter)
    assert _rel(r0.x, single.x) <= 1e-5 and _rel(r0.y, single.y) <= 1e-4
