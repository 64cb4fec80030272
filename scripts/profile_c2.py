#!/usr/bin/env python
"""Config 2 (Lasso n=1e5, m=2e5, 62.5 M nnz): per-phase device times of a bounded solve."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "osqp-python_b200"))
import numpy as np
import osqp
from b200qp import problems
t0 = time.time()
P, q, A, l, u = problems.lasso_qp(25000, 50000, density=0.05, seed=1)
print("gen", round(time.time() - t0, 1), "s", A.shape, A.nnz, flush=True)
prob = osqp.OSQP()
t0 = time.time()
prob.setup(P, q, A, l, u, verbose=False, warm_starting=False, max_iter=int(os.environ.get("C2_ITERS", 25)))
print("setup", round(time.time() - t0, 1), "s", flush=True)
for rep in range(2):
    r = prob.solve()
    st = prob.stats()
    print(json.dumps(dict(status=r.info.status, iters=int(r.info.iter), cg=int(st["cg_iters"]), ms=st["kernel_ms"], grid=st["grid"],
                          phases=dict(zip(["R", "C1", "C2", "C3", "C4", "Z"], [round(v, 2) for v in st["phase_ms"][:6]])),
                          algo_gb=st["algo_bytes"] / 1e9, gbs=st["algo_bytes"] / st["kernel_ms"] / 1e6)), flush=True)
