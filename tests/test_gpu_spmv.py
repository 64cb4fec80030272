"""GPU parity of the hot operator through the raw C ABI (ctypes -> b200qp_spmv_*): every SpMV
variant against scipy on seeded matrices, including the ragged / empty / long-row edge cases."""
import numpy as np
import pytest
from scipy import sparse

pytestmark = pytest.mark.gpu


def _mats():
    rng = np.random.default_rng(0)
    out = {}
    out["uniform10"] = sparse.random(5000, 4000, density=10 / 4000, random_state=rng, format="csr")
    A = sparse.random(3000, 3000, density=0.002, random_state=rng, format="lil")
    A[17, :] = rng.standard_normal(3000)          # one dense row (> one staged tile)
    A[100:140, :] = 0                             # empty rows
    out["ragged"] = A.tocsr()
    out["tiny"] = sparse.csr_matrix(np.array([[1.0, 2.0], [0.0, 0.0], [3.0, 0.0]]))
    out["empty"] = sparse.csr_matrix((7, 5))
    out["dense_rows"] = sparse.random(64, 9000, density=0.6, random_state=rng, format="csr")
    out["wide"] = sparse.random(20000, 50, density=0.2, random_state=rng, format="csr")
    # row lengths 0..200 in no order: exercises the ballot-compacted columns of the tile layout (lanes drop out one by one)
    lens = rng.integers(0, 201, size=1000)
    rows = np.repeat(np.arange(1000), lens)
    cols = np.concatenate([rng.choice(7000, size=k, replace=False) for k in lens])
    out["mixed_lengths"] = sparse.csr_matrix((rng.standard_normal(len(rows)), (rows, cols)), shape=(1000, 7000))
    return out


def _tile_eligible(M):
    """The tile layout stores row lengths in a byte and needs every 32-row tile to fit one ring stage (csl_host.h)."""
    lens = np.diff(M.indptr)
    if lens.size == 0:
        return True
    if lens.max() > 255:
        return False
    pad = (-len(lens)) % 32
    per_tile = np.pad(lens, (0, pad)).reshape(-1, 32).sum(axis=1)
    return 80 + 12 * (int(per_tile.max()) + 3) <= (100 * 1024) // 10


@pytest.mark.parametrize("name", list(_mats().keys()))
def test_spmv_variants_match_scipy(name):
    from b200qp import lowlevel
    M = _mats()[name]
    M.sort_indices()
    x = np.random.default_rng(1).standard_normal(M.shape[1])
    want = M @ x
    scale = max(1.0, float(np.max(np.abs(want))) if want.size else 1.0)
    assert "tile-tma" in lowlevel.variants()
    for v, vname in enumerate(lowlevel.variants()):
        if vname == "tile-tma" and not _tile_eligible(M):
            with pytest.raises(RuntimeError):          # refused loudly, never a silent fallback
                lowlevel.spmv_host(M, x, variant=v)
            continue
        got = lowlevel.spmv_host(M, x, variant=v)
        assert np.max(np.abs(got - want)) <= 1e-12 * scale, f"variant {vname} on {name}"


def test_spmv_plan_linearity_at_scale():
    """Full config-4 shape (1e6 x 1e6, 1e7 nnz): linearity M(ax+by) = aMx + bMy and variant agreement."""
    from b200qp import lowlevel, problems
    _, _, A, _, _ = problems.random_qp(1_000_000, 1_000_000, nnz_per_row=10, seed=3)
    M = A.tocsr()
    plan = lowlevel.SpmvPlan(M)
    rng = np.random.default_rng(5)
    x, y = rng.standard_normal(M.shape[1]), rng.standard_normal(M.shape[1])
    base = None
    for v in range(len(lowlevel.variants())):
        mx, _ = plan.run(x, variant=v)
        my, _ = plan.run(y, variant=v)
        mxy, _ = plan.run(2.0 * x - 3.0 * y, variant=v)
        assert np.max(np.abs(mxy - (2.0 * mx - 3.0 * my))) <= 1e-11 * max(1.0, np.max(np.abs(mxy)))
        if base is None:
            base = mx
            assert np.max(np.abs(mx - M @ x)) <= 1e-11 * np.max(np.abs(mx))
        else:
            assert np.max(np.abs(mx - base)) <= 1e-11 * np.max(np.abs(base))
    plan.close()
