"""One QP partitioned over several B200s (BASELINE.json configs[3]).

The reference has no multi-GPU path (OSQPSettings.device is one int).  Here every rank holds the full
problem, works on a slab of constraint rows and a slice of the variables, and the ranks' persistent ADMM
kernels exchange the gathered vectors directly over NVLink (peer-mapped memory, all-gather by stores fused
into the producing phases, flat flag barrier; csrc/admm_kernel_csl.cuh <true>, CSR fallback
csrc/admm_kernel_dist.cuh).  Two ways to drive it:

  LocalGroup   one process, one host thread, `world` ranks (one GPU each, or several ranks sharing
               a GPU -- used by the tests to exercise the protocol on a single-GPU box).
  DistSolver   one process per GPU (torchrun); torch.distributed (NCCL) is plumbing only: it carries
               the 64-byte cudaIpc handles once and the timing max (and, with the CSR fallback kernel only,
               the y slabs of the result -- the tile kernel leaves all of y on every rank).
"""
from types import SimpleNamespace

import numpy as np


def _prepare(P, q, A, l, u, settings):
    import osqp
    shell = osqp.OSQP(algebra="cuda")
    m, n, P, q, A, l, u = shell._infer_mnpqalu(P=P, q=q, A=A, l=l, u=u)
    ext = shell.ext
    st = ext.OSQPSettings()
    ext.osqp_set_default_settings(st)
    shell.settings = st
    shell.update_settings(**settings)
    q, l, u = (np.ascontiguousarray(v, dtype=np.float64) for v in (q, l, u))
    return ext, st, ext.CSC(P.astype(np.float64)), q, ext.CSC(A.astype(np.float64)), l, u, m, n


def _result(solver, y_full, pc_full):
    info = solver.info
    fields = {k: getattr(info, k) for k in vars(type(info)) if not k.startswith("_")}
    sol = solver.solution
    return SimpleNamespace(x=sol.x, y=y_full, prim_inf_cert=pc_full, dual_inf_cert=sol.dual_inf_cert,
                           info=SimpleNamespace(**fields))


class LocalGroup:
    """All ranks in this process.  devices[i] = CUDA device of rank i (repeats allowed)."""

    def __init__(self, P, q, A, l, u, devices, **settings):
        from osqp.interface import _reraise
        world = len(devices)
        self.world = world
        self.solvers = []
        ext = None
        for rank, dev in enumerate(devices):
            ext, st, Pc, qc, Ac, lc, uc, m, n = _prepare(P, q, A, l, u, dict(settings, device=dev))
            share = sum(1 for d in devices if d == dev)
            self.solvers.append(_reraise(ext.OSQPSolver, Pc, qc, Ac, lc, uc, m, n, st, rank, world, share))
        self.ext, self.m, self.n = ext, m, n
        _reraise(ext.dist_connect_local, self.solvers)

    def solve(self):
        rc = self.ext.dist_solve_group(self.solvers)
        if rc:
            raise RuntimeError(f"multi-GPU solve failed with code {rc}")
        y = np.zeros(self.m)
        pc = np.zeros(self.m)
        for s in self.solvers:
            lo, hi = s.dist_row_range()
            y[lo:hi] = s.solution.y[lo:hi]
            pc[lo:hi] = s.solution.prim_inf_cert[lo:hi]
        return [_result(s, y, pc) for s in self.solvers]

    def stats(self):
        return [s.get_stats() for s in self.solvers]


def gather_slabs(rows, lo, hi, m, world, dist, torch, dev):
    """All-gather per-rank row slabs of m-vectors: rows = k x m array valid on [lo, hi) -> k full vectors.
    Host-side plumbing of the multi-GPU result path (tested under gloo on CPU)."""
    buf = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float64)).to(dev)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    lo_hi = torch.tensor([lo, hi], dtype=torch.int64, device=dev)
    ranges = [torch.empty_like(lo_hi) for _ in range(world)]
    dist.all_gather(ranges, lo_hi)
    out = np.zeros((buf.shape[0], m))
    covered = np.zeros(m, dtype=bool)
    for part, rg in zip(parts, ranges):
        a, b = int(rg[0]), int(rg[1])
        out[:, a:b] = part.cpu().numpy()[:, a:b]
        covered[a:b] = True
    if not covered.all():
        raise RuntimeError("row slabs of the ranks do not cover [0, m)")
    return out


class DistSolver:
    """This process is rank `rank` of `world`; one GPU per process (torchrun)."""

    def __init__(self, P, q, A, l, u, rank, world, device, **settings):
        import torch
        import torch.distributed as dist
        from osqp.interface import _reraise
        self.rank, self.world, self.device = rank, world, device
        ext, st, Pc, qc, Ac, lc, uc, m, n = _prepare(P, q, A, l, u, dict(settings, device=device))
        self.m, self.n = m, n
        self.solver = _reraise(ext.OSQPSolver, Pc, qc, Ac, lc, uc, m, n, st, rank, world, 1)
        mine = torch.frombuffer(bytearray(self.solver.dist_handle()), dtype=torch.uint8).clone()
        dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
        mine = mine.to(dev)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
        _reraise(self.solver.dist_connect_ipc, blob)
        dist.barrier()
        self._torch, self._dist, self._dev = torch, dist, dev

    def update(self, q=None):
        if q is not None:
            self.solver.update_data_vec(q=q, l=None, u=None)

    def solve(self, gather_y=True):
        rc = self.solver.solve()
        if rc:
            raise RuntimeError(f"multi-GPU solve failed with code {rc} on rank {self.rank}")
        y, pc = self.solver.solution.y, self.solver.solution.prim_inf_cert
        lo, hi = self.solver.dist_row_range()
        # tile kernel: the last check iteration pushed y and the certificate to every rank over NVLink -- the row range is
        # [0, m) and there is nothing to gather; the CSR fallback kernel keeps slabs and needs the host-side all-gather
        if gather_y and self.world > 1 and (lo, hi) != (0, self.m):
            y, pc = gather_slabs(np.stack([y, pc]), lo, hi, self.m, self.world, self._dist, self._torch, self._dev)
        return _result(self.solver, y, pc)

    def stats(self):
        return self.solver.get_stats()


def bench(args, rank, world, local, B):
    """bench.py body for N > 1 (launched by torch.distributed.run, one rank per GPU, NCCL)."""
    import json
    import time
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    P, q, A, l, u = B.make_problem()
    t_setup = time.perf_counter()
    solver = DistSolver(P, q, A, l, u, rank, world, local, verbose=False, **B.SOLVE_KW)
    t_setup = time.perf_counter() - t_setup
    n, m = solver.n, solver.m
    rng = np.random.default_rng(1234)
    q_host = [q, q * (1.0 + 1e-12 * rng.standard_normal(n))]

    def step(i):
        solver.update(q=q_host[i & 1])
        return solver.solve(gather_y=True)

    for i in range(args.warmup):
        step(i)
    sampler = B.ClockSampler(gpu=local)
    dist.barrier(); torch.cuda.synchronize()
    sampler.start()
    dev_ms = wall_ms = 0.0
    iters = cg = 0
    bytes_algo = 0.0
    res = None
    phases = [0.0] * 8
    for i in range(args.steps):
        t0 = time.perf_counter()
        res = step(i)
        wall_ms += 1e3 * (time.perf_counter() - t0)
        st = solver.stats()
        dev_ms += st["kernel_ms"]; iters += st["admm_iters"]; cg += st["cg_iters"]; bytes_algo += st["algo_bytes"]
        phases = [a + b for a, b in zip(phases, st["phase_ms"])]
    dist.barrier(); torch.cuda.synchronize()
    clocks = sampler.stop()
    # max over ranks of the device time and of the wall time
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    batch = None
    if not getattr(args, "no_batch", False):
        from b200qp import batchbench
        batch = batchbench.batch_line(args, rank, world, local, dist=dist, torch=torch)
    if rank == 0:
        peak, peak_src = B.measured_peak()
        achieved = bytes_algo / (dev_ms * 1e-3) / 1e9      # whole-job algorithmic bytes (same model as N = 1)
        line = {
            "metric": B.METRIC, "value": iters / (dev_ms * 1e-3), "unit": B.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": B.config_block(),
            "detail": {"admm_iters_per_step": iters // args.steps,
                       "pcg_iters_per_admm_iter": cg / max(iters, 1), "status": res.info.status,
                       "partition": f"{world} row slabs of A + {world} variable slices of P and A'; u, d, t, w2 (and y on check "
                                    "iterations) all-gathered by NVLink stores fused into the producing phases of the persistent "
                                    "kernel, x replicated",
                       "kernel_kind": st.get("kernel_kind"),
                       "phase_ms_per_step_rank0": dict(zip(["R_slice_push_u", "B_Au_push_t", "C_Ku_slice", "V_slice_push_u", "x4", "Z_push_w2",
                                                            "gather_d", "x7"], [round(v / args.steps, 3) for v in phases])),
                       "grid": f"{st['grid']}x{st['block']} per GPU", "setup_s": t_setup},
            "roofline": {"bound": "hbm", "kernel": ("b200qp::admm_kernel_csl<true> (tiles + TMA ring, flat NVLink barrier; persistent, one launch "
                                                    "per rank per solve)" if st.get("kernel_kind") == 3 else
                                                    "b200qp::admm_kernel_dist (CSR fallback; persistent, one launch per rank per solve)"),
                         "achieved": achieved / world, "peak": peak, "unit": "GB/s", "frac": achieved / world / peak,
                         "peak_source": peak_src, "traffic": None,
                         "note": "per-GPU share of the single-GPU algorithmic byte model / max-over-ranks device time"},
            "e2e": {"value": iters / (wall_ms * 1e-3), "unit": B.UNIT, "h2d_bytes_per_step": 8 * n * world,
                    "d2h_bytes_per_step": (8 * (n + m) + 200) * world, "ms_per_step": wall_ms / args.steps,
                    "api": "DistSolver.update(q=host) + solve() -> x, y on every rank's host (complete on every rank: no gather)"},
            "gpu_launches": args.steps * 6 * world,
            "clocks": clocks,
            "cpu_baseline": {"value": None, "unit": B.UNIT, "cores": 0, "kind": "port", "sample": "N = 1 run only"},
        }
        if batch is not None:
            line["batch"] = batch
            line["gpu_launches"] += args.steps * batch["gpu_launches_per_step"] * world
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
