"""
Generate tests/golden/cases/*.npz  (run in the BUILD container only; needs /root/reference).

Each case file is self-contained: problem data (triu-P and A in CSC, q, l, u), the
settings, a list of operations (update / warm_start / solve), and the expected
results from two sources:

  ref_*     the reference's own golden vectors, /root/reference/src/osqp/tests/solutions/*.npz
            (consumed by the reference tests at e.g. basic_test.py:43, update_matrices_test.py:48)
  purepy_*  output of the reference's in-tree solver /root/reference/src/osqppurepy
            (imported here, run with tight tolerances)

The problem recipes restate the fixtures of the reference tests (file:line cited per
case) so that the golden vectors apply; storing the generated data makes the tests
independent of numpy's legacy RNG on the GPU box.  /root/reference is NOT read at
test time.

    python tests/golden/make_golden.py
"""
import json
import os
import sys
import warnings

import numpy as np
from scipy import sparse

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "src"))
warnings.filterwarnings("ignore")
import osqppurepy  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "cases")
SOL = os.path.join(REF, "src/osqp/tests/solutions")
os.makedirs(OUT, exist_ok=True)


def ref_npz(name):
    p = os.path.join(SOL, name + ".npz")
    return dict(np.load(p)) if os.path.exists(p) else {}


def to_triu_csc(P):
    P = sparse.csc_matrix(P)
    if sparse.tril(P, -1).nnz:
        P = sparse.triu(P, format="csc")
    P = sparse.csc_matrix(P)
    P.sort_indices()
    return P


def full(Pt):
    Pt = sparse.csc_matrix(Pt)
    return sparse.csc_matrix(Pt + sparse.triu(Pt, 1).T)


def purepy_run(P_triu, q, A, l, u, settings, ops):
    """Drive osqppurepy through the same operation list.  Setting names are mapped to
    purepy's; P is expanded to full (8c-1)."""
    st = dict(settings)
    pp = {"verbose": False}
    ren = {"polishing": "polish", "warm_starting": "warm_start"}
    for k, v in st.items():
        if k in ("solver_type", "check_dualgap", "cg_max_iter", "cg_tol_fraction", "rho_is_vec"):
            continue
        if k == "check_termination":
            pp[k] = bool(v)  # purepy: bool, every iteration
            continue
        pp[ren.get(k, k)] = v
    pp.setdefault("check_termination", True)
    m = osqppurepy.OSQP()
    Pf, A = full(P_triu), sparse.csc_matrix(A)
    if A.shape[0] == 0:
        return None  # purepy crashes for m == 0 with scaling on new scipy
    m.setup(P=Pf, q=np.asarray(q, float), A=A, l=np.asarray(l, float), u=np.asarray(u, float), **pp)
    Pcur, Acur = P_triu.copy(), A.copy()
    res = None
    for op in ops:
        kind, kw = op[0], op[1]
        if kind == "solve":
            res = m.solve()
        elif kind == "warm_start":
            # purepy omits the cost scale on y: pre-multiply so that the
            # iterate it starts from is the mathematically intended one.
            c = m._model.work.scaling.c if pp.get("scaling", 10) else 1.0
            m.warm_start(x=np.array(kw["x"], float), y=c * np.array(kw["y"], float))
        elif kind == "update":
            vec = {k: np.array(kw[k], float) for k in ("q", "l", "u") if k in kw}
            if vec:
                m.update(**vec)
            if "Px" in kw or "Ax" in kw:
                if "Px" in kw:
                    idx = np.array(kw["Px_idx"]) if "Px_idx" in kw else np.arange(Pcur.nnz)
                    Pcur.data[idx] = np.array(kw["Px"], float)
                if "Ax" in kw:
                    idx = np.array(kw["Ax_idx"]) if "Ax_idx" in kw else np.arange(Acur.nnz)
                    Acur.data[idx] = np.array(kw["Ax"], float)
                # purepy's update_P_A drops c (8c-7): update one at a time instead
                w = m._model
                if "Px" in kw:
                    w.update_P(full(Pcur))
                if "Ax" in kw:
                    w.update_A(Acur)
    info = res.info
    names = {1: "solved", 2: "solved inaccurate", -2: "maximum iterations reached",
             -3: "primal infeasible", -4: "dual infeasible", 3: "primal infeasible inaccurate",
             4: "dual infeasible inaccurate", -7: "problem non convex", -10: "unsolved"}
    out = {"purepy_status": names[info.status_val], "purepy_iter": info.iter,
           "purepy_obj": float(info.obj_val)}
    w = m._model.work
    if info.status_val in (1, 2, -2):
        out["purepy_x"] = np.array(res.x, float)
        out["purepy_y"] = np.array(res.y, float)
        out["purepy_status_polish"] = int(info.status_polish)
    elif info.status_val == -3:
        out["purepy_prim_inf_cert"] = np.array(w.delta_y, float)
    elif info.status_val == -4:
        out["purepy_dual_inf_cert"] = np.array(w.delta_x, float)
    return out


def save(name, P, q, A, l, u, settings, ops, ref=None, note=""):
    P = to_triu_csc(P)
    A = sparse.csc_matrix(A)
    A.sort_indices()
    n, m = P.shape[0], A.shape[0]
    q, l, u = np.asarray(q, float), np.asarray(l, float), np.asarray(u, float)
    d = dict(n=n, m=m, P_data=P.data.astype(float), P_indices=P.indices.astype(np.int32),
             P_indptr=P.indptr.astype(np.int32), A_data=A.data.astype(float),
             A_indices=A.indices.astype(np.int32), A_indptr=A.indptr.astype(np.int32),
             q=q, l=l, u=u, settings=json.dumps(settings), ops=json.dumps(ops), note=note)
    for k, v in (ref or {}).items():
        d["ref_" + k] = v
    pp = purepy_run(P, q, A, l, u, settings, ops)
    if pp:
        d.update(pp)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name:34s} n={n:4d} m={m:4d} purepy={pp and (pp['purepy_status'], pp['purepy_iter'])}")


def L(a):
    return np.asarray(a).tolist()


# ---------------------------------------------------------------- basic_test.py:11-37
def basic():
    P = sparse.diags([11.0, 0.0], format="csc")
    q = np.array([3.0, 4.0])
    A = sparse.csc_matrix([[-1, 0], [0, -1], [-1, -3], [2, 5], [3, 4]], dtype=float)
    u = np.array([0.0, 0.0, -15, 100, 80])
    l = -1e06 * np.ones(len(u))
    opts = {"eps_abs": 1e-09, "eps_rel": 1e-09, "max_iter": 2500, "rho": 0.1, "adaptive_rho": 0,
            "polishing": 0, "check_termination": 1, "warm_starting": 1}
    return P, q, A, l, u, opts


P, q, A, l, u, o = basic()
save("test_basic_QP", P, q, A, l, u, o, [["solve", {}]], ref_npz("test_basic_QP"), "basic_test.py:40-46")
save("test_update_q", P, q, A, l, u, o, [["update", {"q": [10, 20]}], ["solve", {}]],
     ref_npz("test_update_q"), "basic_test.py:49-58")
save("test_update_l", P, q, A, l, u, o, [["update", {"l": L(-50 * np.ones(5))}], ["solve", {}]],
     ref_npz("test_update_l"), "basic_test.py:61-70")
save("test_update_u", P, q, A, l, u, o, [["update", {"u": L(1000 * np.ones(5))}], ["solve", {}]],
     ref_npz("test_update_u"), "basic_test.py:73-82")
save("test_update_bounds", P, q, A, l, u, o,
     [["update", {"l": L(-100 * np.ones(5)), "u": L(1000 * np.ones(5))}], ["solve", {}]],
     ref_npz("test_update_bounds"), "basic_test.py:85-99")


# ------------------------------------------------------- update_matrices_test.py:11-42
def upd_mat():
    np.random.seed(1)
    n, m, p = 5, 8, 0.7
    Pt = sparse.random(n, n, density=p)
    Pt_new = Pt.copy()
    Pt_new.data += 0.1 * np.random.randn(Pt.nnz)
    P = (Pt.T.dot(Pt) + sparse.eye(n)).tocsc()
    P_new = (Pt_new.T.dot(Pt_new) + sparse.eye(n)).tocsc()
    P_triu_new = sparse.triu(P_new)
    q = np.random.randn(n)
    A = sparse.random(m, n, density=p, format="csc")
    A_new = A.copy()
    A_new.data += np.random.randn(A_new.nnz)
    l = np.zeros(m)
    u = 30 + np.random.randn(m)
    return P, q, A, l, u, sparse.csc_matrix(P_triu_new), A_new


P, q, A, l, u, Pn, An = upd_mat()
o = {"eps_abs": 1e-08, "eps_rel": 1e-08}
Pn.sort_indices()
An.sort_indices()
pidx, aidx = L(np.arange(Pn.nnz)), L(np.arange(An.nnz))
save("test_solve", P, q, A, l, u, o, [["solve", {}]], ref_npz("test_solve"), "update_matrices_test.py:45-51")
save("test_update_P", P, q, A, l, u, o, [["update", {"Px": L(Pn.data), "Px_idx": pidx}], ["solve", {}]],
     ref_npz("test_update_P"), "update_matrices_test.py:54-66")
save("test_update_P_allind", P, q, A, l, u, o, [["update", {"Px": L(Pn.data)}], ["solve", {}]],
     ref_npz("test_update_P_allind"), "update_matrices_test.py:69-79")
save("test_update_A", P, q, A, l, u, o, [["update", {"Ax": L(An.data), "Ax_idx": aidx}], ["solve", {}]],
     ref_npz("test_update_A"), "update_matrices_test.py:82-94")
save("test_update_A_allind", P, q, A, l, u, o, [["update", {"Ax": L(An.data)}], ["solve", {}]],
     ref_npz("test_update_A_allind"), "update_matrices_test.py:97-107")
save("test_update_P_A_indP_indA", P, q, A, l, u, o,
     [["update", {"Px": L(Pn.data), "Px_idx": pidx, "Ax": L(An.data), "Ax_idx": aidx}], ["solve", {}]],
     ref_npz("test_update_P_A_indP_indA"), "update_matrices_test.py:110-124")
save("test_update_P_A_indP", P, q, A, l, u, o,
     [["update", {"Px": L(Pn.data), "Px_idx": pidx, "Ax": L(An.data)}], ["solve", {}]],
     ref_npz("test_update_P_A_indP"), "update_matrices_test.py:127-140")
save("test_update_P_A_indA", P, q, A, l, u, o,
     [["update", {"Px": L(Pn.data), "Ax": L(An.data), "Ax_idx": aidx}], ["solve", {}]],
     ref_npz("test_update_P_A_indA"), "update_matrices_test.py:143-156")
save("test_update_P_A_allind", P, q, A, l, u, o,
     [["update", {"Px": L(Pn.data), "Ax": L(An.data)}], ["solve", {}]],
     ref_npz("test_update_P_A_allind"), "update_matrices_test.py:159-170")

# ------------------------------------------------------------ polishing_test.py:11-29
po = {"eps_abs": 1e-03, "eps_rel": 1e-03, "scaling": 10, "rho": 0.1, "alpha": 1.6, "max_iter": 2500,
      "polishing": 1, "polish_refine_iter": 4}
P = sparse.diags([11.0, 0.0], format="csc")
q = np.array([3.0, 4.0])
A = sparse.csc_matrix([[-1, 0], [0, -1], [-1, -3], [2, 5], [3, 4]], dtype=float)
u = np.array([0.0, 0, -15, 100, 80])
l = -1e05 * np.ones(len(u))
save("test_polish_simple", P, q, A, l, u, po, [["solve", {}]], ref_npz("test_polish_simple"),
     "polishing_test.py:32-50")
np.random.seed(4)
n = 30
P = (sparse.diags(np.random.rand(n)) + 0.2 * sparse.eye(n)).tocsc()
q = np.random.randn(n)
save("test_polish_unconstrained", P, q, sparse.csc_matrix((0, n)), np.array([]), np.array([]), po,
     [["solve", {}]], ref_npz("test_polish_unconstrained"), "polishing_test.py:53-73")
np.random.seed(6)
n, m = 30, 50
Pt = sparse.random(n, n)
P = Pt.T @ Pt
q = np.random.randn(n)
A = sparse.csc_matrix(np.random.randn(m, n))
l = -3 + np.random.randn(m)
u = 3 + np.random.randn(m)
save("test_polish_random", P, q, A, l, u, po, [["solve", {}]], ref_npz("test_polish_random"),
     "polishing_test.py:76-99")
# same problem, no polish, tight tolerance: the SURVEY 8c equivalence probe (811 iterations)
save("polish_random_tight", P, q, A, l, u,
     {"eps_abs": 1e-9, "eps_rel": 1e-9, "max_iter": 20000, "polishing": 0, "check_termination": 1,
      "adaptive_rho_interval": 200},
     [["solve", {}]], {}, "polishing_test.py:76-91 problem at eps 1e-9")

# ----------------------------------------------------------- feasibility_test.py:11-43
np.random.seed(4)
n = m = 30
P = sparse.csc_matrix((n, n))
q = np.zeros(n)
A = sparse.random(m, n, density=1.0, format="csc")
u = np.random.rand(m)
save("test_feasibility_problem", P, q, A, u, u,
     {"eps_abs": 1e-06, "eps_rel": 1e-06, "scaling": 10, "alpha": 1.6, "max_iter": 5000, "polishing": 0,
      "warm_starting": 1, "polish_refine_iter": 4},
     [["solve", {}]], ref_npz("test_feasibility_problem"), "feasibility_test.py:46-56")

# --------------------------------------------------------- unconstrained_test.py:11-34
np.random.seed(4)
n = 30
P = (sparse.diags(np.random.rand(n)) + 0.2 * sparse.eye(n)).tocsc()
q = np.random.randn(n)
save("test_unconstrained_problem", P, q, sparse.csc_matrix((0, n)), np.array([]), np.array([]),
     {"eps_abs": 1e-08, "eps_rel": 1e-08, "polishing": 0}, [["solve", {}]],
     ref_npz("test_unconstrained_problem"), "unconstrained_test.py:37-45")

# -------------------------------------------------- primal_infeasibility_test.py:24-58
np.random.seed(4)
n, m = 50, 500
Pt = np.random.rand(n, n)
P = sparse.triu(Pt.T.dot(Pt), format="csc")
q = np.random.rand(n)
A = sparse.random(m, n).tolil()
u = 3 + np.random.randn(m)
l = -3 + np.random.randn(m)
A[int(n / 2), :] = A[int(n / 2) + 1, :]
l[int(n / 2)] = u[int(n / 2) + 1] + 10 * np.random.rand()
u[int(n / 2)] = l[int(n / 2)] + 0.5
A = A.tocsc()
pio = {"eps_abs": 1e-05, "eps_rel": 1e-05, "eps_dual_inf": 1e-20, "max_iter": 2500, "polishing": 0}
save("test_primal_infeasibility", P, q, A, l, u, pio, [["solve", {}]],
     ref_npz("test_primal_infeasibility"), "primal_infeasibility_test.py:24-58")
P2 = sparse.csc_matrix((2, 2))
q2 = np.array([-1.0, -1.0])
A2 = sparse.csc_matrix([[1.0, -1.0], [-1.0, 1.0], [1.0, 0.0], [0.0, 1.0]])
l2 = np.array([1.0, 1.0, 0.0, 0.0])
u2 = np.inf * np.ones(4)
save("test_primal_and_dual_infeasible", P2, q2, A2, l2, np.minimum(u2, 1e30), pio, [["solve", {}]], {},
     "primal_infeasibility_test.py:61-77")

# ---------------------------------------------------- dual_infeasibility_test.py:11-97
dio = {"eps_abs": 1e-05, "eps_rel": 1e-05, "eps_prim_inf": 1e-15, "eps_dual_inf": 1e-6, "scaling": 3,
       "max_iter": 2500, "polishing": 0, "check_termination": 1, "polish_refine_iter": 4}
dref = ref_npz("test_dual_infeasibility")
save("test_dual_infeasible_lp", sparse.csc_matrix((2, 2)), np.array([2.0, -1.0]), sparse.eye(2, format="csc"),
     np.array([0.0, 0.0]), np.array([1e30, 1e30]), dio, [["solve", {}]],
     {"normalized_dual_inf_cert": dref["lp_normalized_dual_inf_cert_correct"]},
     "dual_infeasibility_test.py:31-50")
save("test_dual_infeasible_qp", sparse.diags([4.0, 0.0], format="csc"), np.array([0.0, 2.0]),
     sparse.csc_matrix([[1.0, 1.0], [-1.0, 1.0]]), np.array([-1e30, -1e30]), np.array([2.0, 3.0]), dio,
     [["solve", {}]], {"normalized_dual_inf_cert": dref["qp_normalized_dual_inf_cert_correct"]},
     "dual_infeasibility_test.py:53-72")
save("test_primal_and_dual_infeasible_ws", P2, q2, A2, l2, np.minimum(u2, 1e30), dio,
     [["warm_start", {"x": [25.0, 25.0], "y": [-2.0, -2.0, -2.0, -2.0]}], ["solve", {}]], {},
     "dual_infeasibility_test.py:75-97")

# ------------------------------------------------------------- warm_start_test.py:25-41
np.random.seed(2)
n, m = 100, 200
A = sparse.random(m, n, density=0.9, format="csc")
l = -np.random.rand(m) * 2.0
u = np.random.rand(m) * 2.0
Pr = sparse.random(n, n, density=0.9)
P = sparse.triu(Pr.dot(Pr.T), format="csc")
q = np.random.randn(n)
save("test_warm_start", P, q, A, l, u,
     {"adaptive_rho": 0, "eps_abs": 1e-08, "eps_rel": 1e-08, "polishing": 0, "check_termination": 1},
     [["solve", {}]], {}, "warm_start_test.py:25-57 (direct-solver tolerances)")

# -------------------------------------------------------------- non_convex_test.py:11-20
Pnc = sparse.triu([[2.0, 5.0], [5.0, 1.0]], format="csc")
save("test_non_convex_big_sigma", Pnc, np.array([3.0, 4.0]),
     sparse.csc_matrix([[-1.0, 0.0], [0.0, -1.0], [-1.0, 3.0], [2.0, 5.0], [3.0, 4]]),
     -1e30 * np.ones(5), np.array([0.0, 0.0, -15, 100, 80]), {"sigma": 5}, [["solve", {}]], {},
     "non_convex_test.py:52-56")


# ------------------------------------------ BASELINE config 1: random sparse QP n=100 m=200
def c1_problem(seed=0, n=100, m=200):
    """SURVEY 8d C1 (our generator; duplicated in b200qp/problems.py::random_qp)."""
    rng = np.random.default_rng(seed)
    M = sparse.random(n, n, density=1.0 / n, random_state=rng, data_rvs=rng.standard_normal, format="csc")
    P = (M.T @ M + 1e-2 * sparse.eye(n)).tocsc()
    rows = np.repeat(np.arange(m), 3)
    cols = rng.integers(0, n, size=3 * m)
    A = sparse.csc_matrix((rng.standard_normal(3 * m), (rows, cols)), shape=(m, n))
    A.sum_duplicates()
    x0 = rng.standard_normal(n)
    ax = A @ x0
    sl, su = rng.random(m), rng.random(m)
    kind = rng.random(m)
    l = np.where(kind < 0.15, ax, np.where(kind < 0.20, -1e30, ax - sl))
    u = np.where(kind < 0.15, ax, np.where((kind >= 0.20) & (kind < 0.25), 1e30, ax + su))
    q = rng.standard_normal(n)
    return P, q, A, l, u


P, q, A, l, u = c1_problem()
save("c1_random_qp", P, q, A, l, u,
     {"eps_abs": 1e-9, "eps_rel": 1e-9, "max_iter": 20000, "check_termination": 1,
      "adaptive_rho_interval": 50, "polishing": 0},
     [["solve", {}]], {}, "BASELINE.json configs[0]: n=100 m=200 nnz~600")

# copy the reference's own golden files verbatim (data, not source) for provenance checks
import shutil  # noqa: E402
os.makedirs(os.path.join(HERE, "ref_solutions"), exist_ok=True)
for f in sorted(os.listdir(SOL)):
    if f.endswith(".npz"):
        shutil.copy(os.path.join(SOL, f), os.path.join(HERE, "ref_solutions", f))
print("done")
