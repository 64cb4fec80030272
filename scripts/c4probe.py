"""Config-4 family probe: cold solve at eps 1e-3, device phase times.  usage: c4probe.py [rand|band] [n]"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "osqp-python_b200"))
import numpy as np, osqp
from b200qp import problems
band = 64 if len(sys.argv) > 1 and sys.argv[1] == 'band' else None
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1000000
P, q, A, l, u = problems.random_qp(n, n, nnz_per_row=10, seed=3, band=band)
prob = osqp.OSQP(); prob.setup(P, q, A, l, u, verbose=False, eps_abs=1e-3, eps_rel=1e-3, warm_starting=False)
for rep in range(3):
    r = prob.solve(); st = prob.stats()
cg = int(st['cg_iters']); ph = st['phase_ms']
print(json.dumps(dict(n=n, band=band, status=r.info.status, it=int(r.info.iter), cg=cg, ms=round(st['kernel_ms'], 2),
                      phase=[round(x, 2) for x in ph], grid=st.get('grid'),
                      B_us_per_1e7nnz=round(ph[1] / cg * 1e3 * 1e6 / n, 1), C_us_per_1e7=round(ph[2] / cg * 1e3 * 1e6 / n, 1))))
