#!/usr/bin/env python
"""GPU sweep: ADMM / PCG iteration counts and kernel time vs cg_tol_fraction on configs 2, 3, 4 (cold solves)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "osqp-python_b200"))
import numpy as np
import osqp
from b200qp import problems

def run(name, P, q, A, l, u, fracs, **kw):
    out = []
    for f in fracs:
        prob = osqp.OSQP()
        prob.setup(P, q, A, l, u, verbose=False, warm_starting=False, cg_tol_fraction=f, **kw)
        prob.solve()
        r = prob.solve()
        st = prob.stats()
        k = problems.kkt_report(P, q, A, l, u, r.x, r.y) if r.info.status == "solved" else {}
        out.append(dict(cfg=name, frac=f, status=r.info.status, iters=int(r.info.iter), cg=int(st["cg_iters"]), ms=round(st["kernel_ms"], 2),
                        obj=float(r.info.obj_val), rho_updates=int(r.info.rho_updates), prim=k.get("prim"), dual=k.get("dual")))
        print(json.dumps(out[-1]), flush=True)
        del prob
    return out

fracs = [0.05, 0.1, 0.15, 0.2, 0.3, 0.5]
which = sys.argv[1:] or ["c4", "c2", "c3"]
if "c4" in which:
    run("c4", *problems.random_qp(1_000_000, 1_000_000, nnz_per_row=10, seed=3), fracs)
    run("c4_eps1e-5", *problems.random_qp(1_000_000, 1_000_000, nnz_per_row=10, seed=3), fracs, eps_abs=1e-5, eps_rel=1e-5)
if "c2" in which:
    run("c2", *problems.lasso_qp(25000, 50000, density=0.05, seed=1), fracs)
if "c3" in which:
    run("c3", *problems.svm_qp(10000, 40000, density=0.01, seed=2, box_rows=20000), fracs)
