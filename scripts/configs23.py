#!/usr/bin/env python
"""BASELINE configs[1] (Lasso-as-QP n=1e5, m=2e5) and configs[2] (soft-SVM n=5e4, m=1e5, adaptive rho + warm start +
polish) at full size on one B200: device time of the solve, iteration counts, KKT residuals of the returned point."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "osqp-python_b200"))
import numpy as np
import osqp
from b200qp import problems


def line(name, prob, res, extra=None):
    st = prob.stats()
    out = dict(cfg=name, status=res.info.status, iters=int(res.info.iter), cg=int(st["cg_iters"]), kernel_ms=round(st["kernel_ms"], 2),
               solve_s=round(res.info.solve_time, 4), polish=int(res.info.status_polish), obj=float(res.info.obj_val),
               gbs=round(st["algo_bytes"] / max(st["kernel_ms"], 1e-9) / 1e6, 1), rho_updates=int(res.info.rho_updates))
    out.update(extra or {})
    print(json.dumps(out), flush=True)


t0 = time.time()
P, q, A, l, u = problems.lasso_qp(25000, 50000, density=0.05, seed=1)
prob = osqp.OSQP()
prob.setup(P, q, A, l, u, verbose=False)
print("c2 gen+setup", round(time.time() - t0, 1), "s", A.shape, A.nnz, flush=True)
r = prob.solve()
k = problems.kkt_report(P, q, A, l, u, r.x, r.y)
line("c2_lasso", prob, r, dict(prim=k["prim"], dual=k["dual"]))
del prob

t0 = time.time()
P, q, A, l, u = problems.svm_qp(10000, 40000, density=0.01, seed=2, box_rows=20000)
prob = osqp.OSQP()
prob.setup(P, q, A, l, u, verbose=False, adaptive_rho=True, polishing=True)
print("c3 gen+setup", round(time.time() - t0, 1), "s", A.shape, A.nnz, flush=True)
r = prob.solve()
k = problems.kkt_report(P, q, A, l, u, r.x, r.y)
line("c3_svm_cold_polish", prob, r, dict(prim=k["prim"], dual=k["dual"]))
q2 = q * (1.0 + 0.01 * np.random.default_rng(0).standard_normal(len(q)))
prob.update(q=q2)
prob.warm_start(x=r.x, y=r.y)
r2 = prob.solve()
k = problems.kkt_report(P, q2, A, l, u, r2.x, r2.y)
line("c3_svm_warm_polish", prob, r2, dict(prim=k["prim"], dual=k["dual"]))
