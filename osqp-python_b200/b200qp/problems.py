"""Synthetic QP generators for the five BASELINE.json configurations.

All data fp64, numpy.random.default_rng(seed); P is PSD by construction and the constraint set is
non-empty by construction (l <= A x0 <= u).  Every generator returns (P, q, A, l, u) with P as the
full symmetric CSC matrix (the facade takes the upper triangle itself) and A in CSC.
"""
import numpy as np
from scipy import sparse

INF = 1e30


def _distinct_rows(cols, n):
    """Make the column indices of every row distinct (cols: rows x k, modified copy)."""
    cols = np.sort(cols, axis=1)
    for _ in range(16):
        dup = np.zeros(cols.shape, dtype=bool)
        dup[:, 1:] = cols[:, 1:] == cols[:, :-1]
        if not dup.any():
            break
        cols[dup] = (cols[dup] + 1) % n
        cols = np.sort(cols, axis=1)
    return cols


def random_qp(n, m, nnz_per_row=10, seed=3, eq_frac=0.10, onesided_frac=0.0, p_pairs=1, band=None):
    """Configs 1 and 4: random sparse QP.  A has exactly `nnz_per_row` entries per row (the
    'diagonal' i mod n plus uniform random columns, N(0,1) values); P = diagonally dominant
    symmetric matrix with diag in [0.1, 1.1] plus `p_pairs` symmetric off-diagonal pairs per row.
    band=b keeps every column index within +-b of the row's diagonal position (same n, m, nnz; the
    "sparsity with locality" counterpart of config 4 used by bench.py's structured line)."""
    rng = np.random.default_rng(seed)
    k = min(nnz_per_row, n)
    if band is None:
        cols = rng.integers(0, n, size=(m, k), dtype=np.int64)
    else:
        cols = (np.arange(m, dtype=np.int64)[:, None] % n + rng.integers(-band, band + 1, size=(m, k))) % n
    cols[:, 0] = np.arange(m) % n
    cols = _distinct_rows(cols, n).astype(np.int32)
    vals = rng.standard_normal((m, k))
    indptr = (np.arange(m + 1, dtype=np.int64) * k).astype(np.int32)
    A = sparse.csr_matrix((vals.ravel(), cols.ravel(), indptr), shape=(m, n)).tocsc()
    A.sort_indices()
    # P
    i = np.repeat(np.arange(n), p_pairs)
    j = rng.integers(0, n, size=n * p_pairs) if band is None else (i + rng.integers(-band, band + 1, size=n * p_pairs)) % n
    keep = i != j
    i, j = i[keep], j[keep]
    v = rng.uniform(-0.02, 0.02, size=i.size)
    off = sparse.coo_matrix((np.concatenate([v, v]), (np.concatenate([i, j]), np.concatenate([j, i]))), shape=(n, n)).tocsr()
    off.sum_duplicates()
    rowabs = np.asarray(abs(off).sum(axis=1)).ravel()
    P = (off + sparse.diags(0.1 + rng.random(n) + rowabs)).tocsc()
    P.sort_indices()
    q = rng.standard_normal(n)
    x0 = rng.standard_normal(n)
    ax = A @ x0
    kind = rng.random(m)
    sl, su = rng.random(m), rng.random(m)
    l = np.where(kind < eq_frac, ax, ax - sl)
    u = np.where(kind < eq_frac, ax, ax + su)
    if onesided_frac > 0:
        one = (kind >= eq_frac) & (kind < eq_frac + onesided_frac)
        half = rng.random(m) < 0.5
        l = np.where(one & half, -INF, l)
        u = np.where(one & ~half, INF, u)
    return P, q, A, l, u


def lasso_qp(n_feat, n_data, density=0.05, seed=1, box=True):
    """Config 2: Lasso as a QP (OSQP benchmark form), variables (x, y, t):
         min 1/2 y'y + lam 1't   s.t.  y = Ad x - b,  -t <= x <= t   [+ box rows on every variable]
       n = 2 n_feat + n_data, m = n_data + 2 n_feat (+ n if box).  BASELINE's n=1e5, m=2e5 is
       n_feat=25000, n_data=50000 with the box rows."""
    rng = np.random.default_rng(seed)
    Ad = sparse.random(n_data, n_feat, density=density, random_state=rng, data_rvs=rng.standard_normal, format="csc")
    x_true = rng.standard_normal(n_feat) * (rng.random(n_feat) < 0.5) / np.sqrt(n_feat)
    b = Ad @ x_true + rng.standard_normal(n_data)
    lam = 0.2 * np.max(np.abs(Ad.T @ b))
    n = 2 * n_feat + n_data
    P = sparse.block_diag([sparse.csc_matrix((n_feat, n_feat)), sparse.eye(n_data),
                           sparse.csc_matrix((n_feat, n_feat))], format="csc")
    q = np.concatenate([np.zeros(n_feat + n_data), lam * np.ones(n_feat)])
    In, Zd = sparse.eye(n_feat), sparse.csc_matrix((n_feat, n_data))
    blocks = [sparse.hstack([Ad, -sparse.eye(n_data), sparse.csc_matrix((n_data, n_feat))]),
              sparse.hstack([In, Zd, -In]),
              sparse.hstack([In, Zd, In])]
    l = [b, -INF * np.ones(n_feat), np.zeros(n_feat)]
    u = [b, np.zeros(n_feat), INF * np.ones(n_feat)]
    if box:
        big = 1e3
        blocks.append(sparse.eye(n))
        l.append(-big * np.ones(n))
        u.append(big * np.ones(n))
    A = sparse.vstack(blocks, format="csc")
    A.sort_indices()
    return P, q, A, np.concatenate(l), np.concatenate(u)


def svm_qp(n_feat, n_data, density=0.01, seed=2, box_rows=0):
    """Config 3: soft-margin SVM, variables (x, t):
         min 1/2 x'x + lam 1't   s.t.  t >= diag(b) Ad x + 1,  t >= 0   [+ box rows on leading vars]
       n = n_feat + n_data, m = 2 n_data + box_rows."""
    rng = np.random.default_rng(seed)
    b = np.where(rng.random(n_data) < 0.5, 1.0, -1.0)
    Ad = sparse.random(n_data, n_feat, density=density, random_state=rng, data_rvs=rng.standard_normal, format="csr")
    shift = sparse.csr_matrix(Ad)
    shift.data = shift.data + 0.3 * np.repeat(b, np.diff(Ad.indptr))     # class-dependent mean
    Ad = sparse.csc_matrix(shift)
    lam = 1.0
    n = n_feat + n_data
    P = sparse.block_diag([sparse.eye(n_feat), sparse.csc_matrix((n_data, n_data))], format="csc")
    q = np.concatenate([np.zeros(n_feat), lam * np.ones(n_data)])
    blocks = [sparse.hstack([sparse.diags(b) @ Ad, -sparse.eye(n_data)]),
              sparse.hstack([sparse.csc_matrix((n_data, n_feat)), sparse.eye(n_data)])]
    l = [-INF * np.ones(n_data), np.zeros(n_data)]
    u = [-np.ones(n_data), INF * np.ones(n_data)]
    if box_rows:
        blocks.append(sparse.eye(box_rows, n))
        l.append(-1e3 * np.ones(box_rows))
        u.append(1e3 * np.ones(box_rows))
    A = sparse.vstack(blocks, format="csc")
    A.sort_indices()
    return P, q, A, np.concatenate(l), np.concatenate(u)


def mpc_qp(seed=4, horizon=10, nx=8, nu=4, shared_seed=4):
    """Config 5: one MPC QP, n = horizon (nx+nu) = 120, m = horizon nx + n + ... = 240:
       variables (x_1..x_N, u_0..u_{N-1}); dynamics equalities x_{k+1} = Ad x_k + Bd u_k (N nx rows),
       box rows on every variable but the last state block pads to m = 240 (N nx + N (nx+nu) - ... )."""
    rs = np.random.default_rng(shared_seed)           # shared dynamics
    Adyn = rs.standard_normal((nx, nx))
    Adyn = 0.9 * Adyn / max(abs(np.linalg.eigvals(Adyn)))
    Bdyn = rs.standard_normal((nx, nu))
    rng = np.random.default_rng(seed)                 # per-QP initial state
    x_init = rng.standard_normal(nx)
    N = horizon
    nX, nU = N * nx, N * nu
    n = nX + nU
    Q, R = sparse.eye(nx), 0.1 * sparse.eye(nu)
    P = sparse.block_diag([sparse.kron(sparse.eye(N), Q), sparse.kron(sparse.eye(N), R)], format="csc")
    q = np.zeros(n)
    # dynamics: x_{k+1} - Ad x_k - Bd u_k = 0  (k = 0: x_1 - Bd u_0 = Ad x_init)
    Ax = sparse.eye(nX) - sparse.kron(sparse.eye(N, k=-1), Adyn)
    Bu = -sparse.kron(sparse.eye(N), Bdyn)
    Aeq = sparse.hstack([Ax, Bu])
    beq = np.zeros(nX)
    beq[:nx] = Adyn @ x_init
    # boxes: |x| <= 5 on all states, |u| <= 1 on all inputs, plus 40 rate rows u_k - u_{k-1} to reach m = 240
    Abox = sparse.eye(n)
    lb = np.concatenate([-5 * np.ones(nX), -np.ones(nU)])
    ub = -lb
    rate = sparse.hstack([sparse.csc_matrix((nU, nX)), sparse.eye(nU) - sparse.eye(nU, k=-nu)])
    blocks, l, u = [Aeq, Abox, rate], [beq, lb, -0.5 * np.ones(nU)], [beq, ub, 0.5 * np.ones(nU)]
    A = sparse.vstack(blocks, format="csc")
    A.sort_indices()
    return P, q, A, np.concatenate(l), np.concatenate(u)


def mpc_batch(batch, seed0=4, horizon=10, nx=8, nu=4, shared_seed=4):
    """Config 5: `batch` MPC QPs (seeds seed0 .. seed0+batch-1) that share P, q, A and differ in the initial
    state, i.e. in the first nx entries of l and u.  Returns (P, q, A, l, u) with l, u of shape (batch, m);
    element b equals mpc_qp(seed=seed0 + b)."""
    P, q, A, l0, u0 = mpc_qp(seed=seed0, horizon=horizon, nx=nx, nu=nu, shared_seed=shared_seed)
    rs = np.random.default_rng(shared_seed)
    Adyn = rs.standard_normal((nx, nx))
    Adyn = 0.9 * Adyn / max(abs(np.linalg.eigvals(Adyn)))
    l = np.tile(l0, (batch, 1))
    u = np.tile(u0, (batch, 1))
    for b in range(batch):
        beq = Adyn @ np.random.default_rng(seed0 + b).standard_normal(nx)
        l[b, :nx] = beq
        u[b, :nx] = beq
    return P, q, A, l, u


def kkt_report(P, q, A, l, u, x, y):
    """Size-independent optimality check of a returned (x, y): residuals of the unscaled KKT system."""
    Pf = sparse.csc_matrix(P)
    if sparse.tril(Pf, -1).nnz == 0 and sparse.triu(Pf, 1).nnz:
        Pf = Pf + sparse.triu(Pf, 1).T
    ax = A @ x
    prim = float(np.max(np.maximum(np.maximum(l - ax, ax - u), 0))) if len(l) else 0.0
    dual = float(np.max(np.abs(Pf @ x + q + A.T @ y)))
    yp, ym = np.maximum(y, 0), np.minimum(y, 0)
    fin_u, fin_l = u < 1e26, l > -1e26
    comp = float(np.max(np.abs(np.where(fin_u, yp * (u - ax), 0)) + np.abs(np.where(fin_l, ym * (l - ax), 0)))) if len(l) else 0.0
    obj = float(0.5 * x @ (Pf @ x) + q @ x)
    return {"prim": prim, "dual": dual, "comp": comp, "obj": obj}
