"""bench.py's config-5 line (BASELINE.json configs[4]): 8192 independent MPC QPs (n = 120, m = 240), sharded
over the ranks with no collective on the data path (SURVEY.md 8e "replicas only").  One step = every QP of the
batch solved cold to eps_abs = eps_rel = 1e-4 after an update(l, u) with new initial states."""
import os
import time

import numpy as np

BATCH = int(os.environ.get("B200QP_BENCH_BATCH", 8192))
EPS = 1e-4
KW = dict(eps_abs=EPS, eps_rel=EPS, max_iter=4000, warm_starting=False)


def shard_range(batch, rank, world):
    """QPs [lo, hi) of rank `rank`: contiguous, sizes differ by at most one, no collective needed."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_counts(dev_ms, wall_ms, counts, dist=None, torch=None, device=None):
    """Whole-job numbers from per-rank ones: time = max over ranks, counters = sum over ranks."""
    if dist is None:
        return dev_ms, wall_ms, np.asarray(counts, dtype=np.float64)
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c = torch.tensor(np.asarray(counts, dtype=np.float64), dtype=torch.float64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t[0]), float(t[1]), c.cpu().numpy()


def batch_line(args, rank=0, world=1, local=0, dist=None, torch=None, cpu_sample=None):
    """cpu_sample: optional callable (P, q, A, l, u, KW, BATCH) -> cpu_baseline dict, supplied by bench.py (the CPU leg
    lives outside the package: nothing here imports the oracle)."""
    import osqp
    from b200qp import problems
    lo, hi = shard_range(BATCH, rank, world)
    P, q, A, l_all, u_all = problems.mpc_batch(BATCH, seed0=4)
    _, _, _, l_alt, u_alt = problems.mpc_batch(BATCH, seed0=4 + BATCH)
    sets = [(np.ascontiguousarray(l_all[lo:hi]), np.ascontiguousarray(u_all[lo:hi])),
            (np.ascontiguousarray(l_alt[lo:hi]), np.ascontiguousarray(u_alt[lo:hi]))]
    prob = osqp.OSQPBatch()
    t_setup = time.perf_counter()
    prob.setup(P, q, A, sets[0][0], sets[0][1], verbose=False, device=local, **KW)
    t_setup = time.perf_counter() - t_setup

    def step(i):
        prob.update_settings(rho=0.1)
        prob.update(l=sets[i & 1][0], u=sets[i & 1][1])     # H2D 2 * 8 * m bytes per QP
        return prob.solve()                                  # one kernel + D2H of x, y, info

    def sync():
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    sync()
    dev_ms = wall_ms = 0.0
    iters = cg = solved = 0
    for i in range(args.steps):
        t0 = time.perf_counter()
        res = step(i)
        wall_ms += 1e3 * (time.perf_counter() - t0)
        st = prob.stats()
        dev_ms += st["kernel_ms"]; iters += st["admm_iters"]; cg += st["cg_iters"]
        solved += int((res.info.status_val == 1).sum())
    sync()
    device = torch.device("cuda", local) if dist is not None else None
    dev_ms, wall_ms, tot = reduce_counts(dev_ms, wall_ms, [iters, cg, solved], dist, torch, device)
    n, m = prob.n, prob.m
    nnz = int(sparse_nnz(P) + 2 * A.nnz)
    line = {
        "metric": "batched_qps_per_sec", "unit": "QPs/s", "value": BATCH * args.steps / (dev_ms * 1e-3),
        "workload": f"BASELINE configs[4]: {BATCH} independent MPC QPs n={n} m={m} (horizon 10, nx=8, nu=4; shared "
                    f"P, q, A; per-QP initial state in l, u), seeds 4..; step = update(l,u) + cold solve of all to eps {EPS}",
        "n_gpus": world, "scaling": "strong", "qps_per_gpu": hi - lo, "ms_per_step": dev_ms / args.steps,
        "admm_iters_per_qp": tot[0] / (BATCH * args.steps), "pcg_iters_per_admm_iter": tot[1] / max(tot[0], 1),
        "admm_iters_per_sec": tot[0] / (dev_ms * 1e-3), "solved_fraction": tot[2] / (BATCH * args.steps),
        "kernel": "b200qp::admm_batch_kernel (one CTA per QP, vectors in shared memory, one launch per step)",
        "grid": f"{st['grid']}x{st['block']}", "smem_bytes_per_cta": st["smem_bytes"], "setup_s": t_setup,
        "e2e": {"value": BATCH * args.steps / (wall_ms * 1e-3), "unit": "QPs/s", "ms_per_step": wall_ms / args.steps,
                "h2d_bytes_per_step": 16 * m * BATCH, "d2h_bytes_per_step": (8 * (n + m) + 64) * BATCH,
                "api": "osqp.OSQPBatch.update(l=host, u=host) + solve() -> x, y, info on host"},
        "gpu_launches_per_step": 4,
        "note": "latency-bound, not HBM-bound: per QP the kernel reads 16m + 8n bytes and writes 16(n+m) bytes once; "
                f"the shared matrices ({12 * nnz} bytes) stay in L1",
    }
    if cpu_sample is not None and rank == 0:
        line["cpu_baseline"] = cpu_sample(P, q, A, l_all, u_all, KW, BATCH)        # all host cores
    return line


def sparse_nnz(P):
    from scipy import sparse
    Pf = sparse.csc_matrix(P)
    return Pf.nnz
