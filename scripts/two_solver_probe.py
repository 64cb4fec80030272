"""bench.py's sequence in small: the random config-4 solver (kept alive), then the banded one in the same process."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "osqp-python_b200")]
import numpy as np
import osqp
from b200qp import problems
n = int(os.environ.get("N", 1000000))
KW = dict(eps_abs=1e-3, eps_rel=1e-3, max_iter=4000, warm_starting=False, polishing=False)
def run(tag, band, reps, keep):
    P, q, A, l, u = problems.random_qp(n, n, nnz_per_row=10, seed=3, band=band)
    prob = osqp.OSQP(algebra="cuda")
    t = time.perf_counter()
    prob.setup(P, q, A, l, u, verbose=False, **KW)
    print(tag, "setup %.2f s" % (time.perf_counter() - t), flush=True)
    for i in range(reps):
        t = time.perf_counter()
        if band is None:
            prob.update(q=q * (1.0 + 1e-12 * (i & 1)))
        res = prob.solve()
        st = prob.stats()
        print(tag, i, res.info.status, res.info.iter, "kernel_ms %.2f" % st["kernel_ms"], "wall %.2f" % (time.perf_counter() - t), "ring", st["ring_stages"], st["ring_stage_bytes"],
              "autotune", st.get("autotune_ms"), flush=True)
    keep.append(prob)
keep = []
order = os.environ.get("ORDER", "random,banded").split(",")
for o in order:
    run(o, None if o == "random" else 64, int(os.environ.get("REPS", 6)), keep)
