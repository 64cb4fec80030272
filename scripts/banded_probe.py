"""Config 4's banded counterpart alone (bench.py's `structured` line): setup + a few cold solves, with stats."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "osqp-python_b200")]
import osqp
from b200qp import problems
n = int(os.environ.get("N", 1000000))
P, q, A, l, u = problems.random_qp(n, n, nnz_per_row=10, seed=3, band=64)
prob = osqp.OSQP(algebra="cuda")
prob.setup(P, q, A, l, u, verbose=False, eps_abs=1e-3, eps_rel=1e-3, max_iter=4000, warm_starting=False, polishing=False)
for i in range(int(os.environ.get("REPS", 4))):
    t = time.perf_counter()
    try:
        res = prob.solve()
    except Exception as e:      # noqa: BLE001
        print("solve", i, "failed:", e, flush=True)
        break
    st = prob.stats()
    print("solve", i, res.info.status, res.info.iter, "cg", st["cg_iters"], "kernel_ms %.2f" % st["kernel_ms"], "wall %.2f" % (time.perf_counter() - t),
          "ring", st["ring_stages"], st["ring_stage_bytes"], "grid", st["grid"], "autotune", st.get("autotune_ms"), flush=True)
