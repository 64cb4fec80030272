"""Polish accuracy probe: engine (GPU) vs oracle direct polish on the polish goldens and random QPs."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "osqp-python_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, osqp
from b200qp import problems
from oracle.osqp_oracle import OracleSolver
from case_runner import load_case

def run(name, P, q, A, l, u, **st):
    st = dict(st); st.setdefault("polishing", 1)
    o = OracleSolver(P, q, A, l, u, linsys="direct", **{k: v for k, v in st.items() if k != "verbose"})
    t0 = time.time(); ro = o.solve(); to = time.time() - t0
    prob = osqp.OSQP(); prob.setup(P, q, A, l, u, verbose=False, **st)
    t0 = time.time(); r = prob.solve(); tg = time.time() - t0
    sc = max(1.0, np.max(np.abs(ro.x)))
    ex = np.max(np.abs(r.x - ro.x)) / sc
    ey = np.max(np.abs(r.y - ro.y)) / max(1.0, np.max(np.abs(ro.y))) if len(ro.y) else 0.0
    print(json.dumps(dict(case=name, n=len(q), m=A.shape[0], oracle_polish=int(ro.info.status_polish), gpu_polish=int(r.info.status_polish),
                          it_o=int(ro.info.iter), it_g=int(r.info.iter), err_x=float(ex), err_y=float(ey),
                          obj_o=float(ro.info.obj_val), obj_g=float(r.info.obj_val), pres_g=float(r.info.prim_res), dres_g=float(r.info.dual_res),
                          pres_o=float(ro.info.prim_res), dres_o=float(ro.info.dual_res), polish_s=float(r.info.polish_time), t_oracle=round(to, 2))), flush=True)

for nm in ["test_polish_simple", "test_polish_unconstrained", "test_polish_random", "polish_random_tight"]:
    c = load_case(nm)
    st = {k: v for k, v in c["settings"].items()}
    run(nm, c["P"], c["q"], c["A"], c["l"], c["u"], **st)
for (n, m, seed) in [(200, 300, 1), (1000, 1500, 2), (3000, 3000, 3)]:
    P, q, A, l, u = problems.random_qp(n, m, nnz_per_row=min(10, n // 4), seed=seed)
    run(f"random_qp_{n}x{m}", P, q, A, l, u)
P, q, A, l, u = problems.svm_qp(500, 2000, density=0.02, seed=2, box_rows=1000)
run("svm_small", P, q, A, l, u, adaptive_rho=1)
P, q, A, l, u = problems.lasso_qp(300, 600, density=0.05, seed=1)
run("lasso_small", P, q, A, l, u)
