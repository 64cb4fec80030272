"""BASELINE config 3 (soft-SVM n = 5e4, m = 1e5) at full size: status_polish at the default and at tight tolerances."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "osqp-python_b200")]
import osqp
from b200qp import problems
P, q, A, l, u = problems.svm_qp(10000, 40000, density=0.01, seed=2, box_rows=20000)
for eps in (1e-3, 1e-5, 1e-7):
    prob = osqp.OSQP()
    prob.setup(P, q, A, l, u, verbose=False, adaptive_rho=True, polishing=True, eps_abs=eps, eps_rel=eps, max_iter=20000)
    t = time.perf_counter()
    r = prob.solve()
    k = problems.kkt_report(P, q, A, l, u, r.x, r.y)
    print("c3 eps %g: %s iters %d status_polish %d obj %.9g prim %.2e dual %.2e solve %.3f s polish %.3f s" % (
        eps, r.info.status, r.info.iter, r.info.status_polish, r.info.obj_val, k["prim"], k["dual"], time.perf_counter() - t, r.info.polish_time), flush=True)
