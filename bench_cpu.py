"""CPU legs of bench.py: the oracle (port of the reference algorithm) timed on the host cores.

Test / measurement infrastructure only -- this is the one place besides tests/ and __graft_entry__.smoke() that
executes oracle/ (see the header of oracle/osqp_oracle.py); nothing under osqp-python_b200/ imports it.
"""
import os
import time

ENGINE_SPEC = dict(cg_max_iter=100, cg_tol_fraction=0.1, rho_eq_factor=10.0)   # the engine's defaults, for the CPU port


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm must use every host core it can (r1: the N >= 2
    reference lines ran single-threaded).  Must be called before the OpenMP helper is loaded."""
    n = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(n)
    from oracle import osqp_oracle
    lib = osqp_oracle._omp_lib()
    if lib and hasattr(lib, "oracle_set_threads"):
        lib.oracle_set_threads(n)
    return n


def config4_sample(P, q, A, l, u, solve_kw, iters, budget_s=None):
    """`iters` ADMM iterations of the cold-start solve in ONE solve() call (engine-spec PCG mode: a direct factorisation
    cannot finish at n = m = 1e6 on random sparsity).  Returns (iterations done, PCG iterations, seconds,
    threads)."""
    from oracle import osqp_oracle
    from oracle.osqp_oracle import OracleSolver
    use_all_host_threads()
    kw = {k: v for k, v in solve_kw.items() if k not in ("warm_starting", "polishing")}
    kw["max_iter"] = int(iters)
    s = OracleSolver(P, q, A, l, u, linsys="indirect", check_termination=0, adaptive_rho=0, **ENGINE_SPEC, **kw)
    t0 = time.perf_counter()
    r = s.solve(max_wall=budget_s)
    dt = time.perf_counter() - t0
    return int(r.info.iter), int(r.info.cg_iters), dt, osqp_oracle.host_threads()


# ---------------------------------------------------------------------------------------------- config 5 (batched MPC)
def _mpc_worker(job):
    """One host process: the reference's way on one core -- one setup, then update(l, u) + cold solve per QP (oracle
    port, direct KKT factorisation as in its qdldl path).  Returns (QPs, iterations, seconds)."""
    import sys
    root = os.path.dirname(os.path.abspath(__file__))
    for p_ in (root, os.path.join(root, "osqp-python_b200")):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    from oracle.osqp_oracle import OracleSolver
    P, q, A, l, u, kw = job
    s = OracleSolver(P, q, A, l[0], u[0], linsys="direct", **{k: v for k, v in kw.items() if k != "warm_starting"})
    t0 = time.perf_counter()
    iters = 0
    for b in range(len(l)):
        s.update(l=l[b], u=u[b])
        s.x[:] = 0; s.z[:] = 0; s.y[:] = 0
        s.update_rho(0.1)
        iters += s.solve().info.iter
    return len(l), iters, time.perf_counter() - t0


def mpc_batch_sample(P, q, A, l, u, kw, batch, per_proc=16, procs=None):
    """CPU baseline of config 5: independent QPs spread over one spawned process per host core, each running the oracle
    port; throughput = QPs / slowest worker's time."""
    procs = procs or min(host_cores(), 32)
    jobs = [(P, q, A, l[i * per_proc:(i + 1) * per_proc], u[i * per_proc:(i + 1) * per_proc], kw) for i in range(procs)]
    jobs = [j for j in jobs if len(j[3])]
    used, res = 1, None
    if len(jobs) > 1:
        try:
            import multiprocessing as mp
            from concurrent.futures import ProcessPoolExecutor
            with ProcessPoolExecutor(len(jobs), mp_context=mp.get_context("spawn")) as pool:
                res = list(pool.map(_mpc_worker, jobs, timeout=120))
            used = len(jobs)
        except Exception:      # noqa: BLE001
            res = None
    if res is None:
        res = [_mpc_worker(jobs[0])]
        used = 1
    count = sum(r[0] for r in res)
    iters = sum(r[1] for r in res)
    dt = max(r[2] for r in res)
    return {"value": count / dt, "unit": "QPs/s", "cores": used, "kind": "port",
            "sample": f"{count} of the {batch} QPs ({iters} ADMM iterations, slowest worker {dt:.2f} s) on {used} host process(es), "
                      f"{per_proc} QPs each: one setup, update(l,u) + cold solve per QP; oracle port, sparse LU of the KKT "
                      f"matrix; {host_cores()} host cores"}
