#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 OSQP ADMM engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[3], random sparse QP n = m = 1e6, nnz(A) = 1e7
(b200qp.problems.random_qp, seed 3) -- the configuration the metric is quoted on; it fits one GPU.
One *step* = one complete cold-start solve() of that QP to the default tolerances
(eps_abs = eps_rel = 1e-3); every step does identical work (the engine is deterministic).

  value        ADMM iterations / second, device-timed: CUDA events around the persistent ADMM kernel
               on the stream it is launched on (b200qp_get_stats().kernel_ms), problem resident in HBM.
  e2e          the same metric through the reference-facing API with HOST buffers: osqp.OSQP.update(q=...)
               (H2D) + solve() + x/y read-back (D2H), wall clock around the calls.
  roofline     the dominant (only hot) kernel is the persistent ADMM kernel: algorithmic HBM bytes of
               the launch (byte model in DESIGN.md, computed by the engine from its own iteration
               counters) / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
  structured   config 4's sizes with banded column indices (same kernels and byte model): what the engine reaches
               when the gathers have locality; sub-object with its own roofline and SpMV figures.
  batch        BASELINE configs[4] as a sub-object: 8192 independent MPC QPs sharded over the GPUs (QPs/s,
               device-timed and end to end; b200qp/batchbench.py).
  cpu_baseline the CPU oracle (port of the reference algorithm, engine-spec PCG mode because a direct
               factorisation cannot finish at this size -- SURVEY.md 8d) on a bounded sample.
  --impl reference   times that CPU implementation as its own arm (rank 0 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "osqp-python_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "admm_iters_per_sec"
UNIT = "ADMM iters/s"
N_VARS = int(os.environ.get("B200QP_BENCH_N", 1_000_000))
N_CONS = int(os.environ.get("B200QP_BENCH_M", N_VARS))
NNZ_ROW = 10
SEED = 3
ENGINE_SPEC = dict(cg_max_iter=100, cg_tol_fraction=0.1, rho_eq_factor=10.0)   # the engine's defaults, for the CPU port
# dram__bytes_read.sum + dram__bytes_write.sum of one admm_kernel launch of this workload, from the committed
# `ncu --set full` capture (231.88 GB + 12.23 GB; the n- and m-vectors stay L2-resident, so it is below the
# algorithmic byte count); only valid for the default problem size
NCU_TRAFFIC_BYTES = 244.110e9 if (N_VARS, N_CONS) == (1_000_000, 1_000_000) else None
NCU_TRAFFIC_SOURCE = "profiles/r1d_admm_kernel_ncu_raw.csv (75 ADMM / 642 PCG iterations per launch)"
SOLVE_KW = dict(eps_abs=1e-3, eps_rel=1e-3, max_iter=4000, warm_starting=False, polishing=False)


def workload_name():
    return (f"BASELINE configs[3]: random sparse QP n={N_VARS} m={N_CONS} nnz(A)={N_CONS * NNZ_ROW} "
            f"(10/row, uniform random columns), nnz(P_full)~{3 * N_VARS}, 10% equality rows, seed {SEED}; "
            f"step = one cold-start solve() to eps_abs=eps_rel=1e-3")


def make_problem():
    from b200qp import problems
    return problems.random_qp(N_VARS, N_CONS, nnz_per_row=NNZ_ROW, seed=SEED)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.rows, self.proc, self.gpu = [], None, gpu

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def dist_env():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    return rank, world, local


# ----------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's CPU implementation of the path on the host cores.  The reference's compiled solver
    (osqp v1.0.0 / QDLDL) is not vendored and osqppurepy cannot travel to the GPU box, so this is the
    oracle port of it; direct factorisation cannot finish at n = m = 1e6 on random sparsity,
    hence the engine-spec PCG mode.  One step = one ADMM iteration (bounded sample), continuing the solve."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle.osqp_oracle import OracleSolver
    P, q, A, l, u = make_problem()
    kw = dict(SOLVE_KW)
    kw.pop("warm_starting"); kw.pop("polishing")
    s = OracleSolver(P, q, A, l, u, linsys="indirect", warm_starting=1, check_termination=0,
                     adaptive_rho=0, **ENGINE_SPEC, **kw)
    cap_s = float(os.environ.get("B200QP_REF_BUDGET_S", 150))
    t_begin = time.perf_counter()
    s.s.max_iter = 1
    done_w = 0
    for _ in range(args.warmup):
        if time.perf_counter() - t_begin > cap_s / 3:
            break
        s.solve(); done_w += 1
    t0 = time.perf_counter()
    iters, cg = 0, 0
    for _ in range(args.steps):
        if time.perf_counter() - t_begin > cap_s:
            break
        r = s.solve()
        iters += 1; cg += r.info.cg_iters
    dt = time.perf_counter() - t0
    val = iters / dt if dt > 0 and iters else 0.0
    from oracle import osqp_oracle
    cores = osqp_oracle.host_threads()
    sample = (f"{iters} ADMM iterations ({cg} PCG iterations) of the same problem after {done_w} warm-up "
              f"iterations; CSR SpMV on {cores} thread(s) (OpenMP helper if faster than scipy on this host) + numpy, "
              f"{os.cpu_count()} host cores")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": iters, "warmup": done_w, "ms_per_step": 1e3 * dt / max(iters, 1), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(), "note": "step = 1 ADMM iteration of the CPU port (bounded sample)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------- our arm
def cpu_baseline_sample(P, q, A, l, u, budget_iters=3):
    from oracle.osqp_oracle import OracleSolver
    kw = dict(SOLVE_KW)
    kw.pop("warm_starting"); kw.pop("polishing")
    kw["max_iter"] = budget_iters
    s = OracleSolver(P, q, A, l, u, linsys="indirect", check_termination=0, adaptive_rho=0, **ENGINE_SPEC, **kw)
    t0 = time.perf_counter()
    r = s.solve()
    dt = time.perf_counter() - t0
    from oracle import osqp_oracle
    cores = osqp_oracle.host_threads()
    return {"value": budget_iters / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {budget_iters} ADMM iterations ({r.info.cg_iters} PCG iterations, {dt:.1f} s) of the same "
                      f"cold-start solve; oracle port, CSR SpMV on {cores} thread(s) (OpenMP helper if faster than scipy "
                      f"on this host) + numpy, {os.cpu_count()} host cores",
            "pcg_iters_per_admm_iter": r.info.cg_iters / budget_iters}


def spmv_line(A, peak):
    """Standalone per-iteration SpMV (y = A x) achieved GB/s, every variant, L2 flushed between launches."""
    from b200qp import lowlevel
    M = A.tocsr()
    plan = lowlevel.SpmvPlan(M)
    x = np.random.default_rng(7).standard_normal(M.shape[1])
    out = {}
    for v, name in enumerate(lowlevel.variants()):
        plan.run(x, variant=v, reps=3, want_y=False)
        _, ms_cold = plan.run(None, variant=v, reps=10, flush_l2=True, want_y=False)
        _, ms_warm = plan.run(None, variant=v, reps=20, flush_l2=False, want_y=False)
        out[name] = {"ms_flushed": ms_cold, "gbs_flushed": plan.algo_bytes() / ms_cold / 1e6,
                     "ms_back_to_back": ms_warm, "gbs_back_to_back": plan.algo_bytes() / ms_warm / 1e6}
    best = max(out, key=lambda k: out[k]["gbs_flushed"])
    res = {"matrix": f"A {M.shape[0]}x{M.shape[1]} nnz={M.nnz}", "algo_bytes": plan.algo_bytes(), "variants": out,
           "best": best, "frac_of_measured_peak": out[best]["gbs_flushed"] / peak,
           "frac_of_8TBs": out[best]["gbs_flushed"] / 8000.0}
    plan.close()
    return res


def structured_line(args, peak):
    """Config 4's sizes (n = m = 1e6, 10 nnz per row of A) with LOCALITY in the column indices (every index within
    +-64 of the diagonal position): the same kernels, the same byte model -- shows what the engine reaches when
    the gathers are not uniformly random (config 4 proper is bound by the L1TEX / L2 gather rate)."""
    import osqp
    from b200qp import problems
    P, q, A, l, u = problems.random_qp(N_VARS, N_CONS, nnz_per_row=NNZ_ROW, seed=SEED, band=64)
    prob = osqp.OSQP(algebra="cuda")
    prob.setup(P, q, A, l, u, verbose=False, **SOLVE_KW)
    for _ in range(2):
        prob.solve()
    dev_ms, iters, cg, bytes_algo = 0.0, 0, 0, 0.0
    steps = max(3, min(args.steps, 5))
    for _ in range(steps):
        res = prob.solve()
        st = prob.stats()
        dev_ms += st["kernel_ms"]; iters += st["admm_iters"]; cg += st["cg_iters"]; bytes_algo += st["algo_bytes"]
    achieved = bytes_algo / (dev_ms * 1e-3) / 1e9
    out = {"workload": f"random sparse QP n={N_VARS} m={N_CONS} nnz(A)={N_CONS * NNZ_ROW}, column indices within +-64 of the "
                       f"diagonal (banded), seed {SEED}; cold-start solve() to eps 1e-3",
           "value": iters / (dev_ms * 1e-3), "unit": UNIT, "ms_per_step": dev_ms / steps, "steps": steps,
           "admm_iters_per_step": iters // steps, "pcg_iters_per_admm_iter": cg / max(iters, 1), "status": res.info.status,
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "frac_of_8TBs": achieved / 8000.0}}
    if not args.no_spmv:
        out["spmv"] = spmv_line(A, peak)
    return out


def run_ours(args):
    rank, world, local = dist_env()
    if world > 1:
        from b200qp import multigpu
        return multigpu.bench(args, rank, world, local, sys.modules[__name__])
    import osqp
    t_gen = time.perf_counter()
    P, q, A, l, u = make_problem()
    t_gen = time.perf_counter() - t_gen
    prob = osqp.OSQP(algebra="cuda")
    t_setup = time.perf_counter()
    prob.setup(P, q, A, l, u, verbose=False, **SOLVE_KW)
    t_setup = time.perf_counter() - t_setup
    n, m = prob.n, prob.m
    rng = np.random.default_rng(1234)
    q_host = [q, q * (1.0 + 1e-12 * rng.standard_normal(n))]   # alternate: forces a real H2D each step

    for i in range(args.warmup):
        prob.update(q=q_host[i & 1])
        prob.solve()
    sampler = ClockSampler(gpu=local)
    sampler.start()
    dev_ms, wall_ms, iters, cg, bytes_algo, launches = 0.0, 0.0, 0, 0, 0.0, 0
    status = None
    phases = [0.0] * 8
    for i in range(args.steps):
        t0 = time.perf_counter()
        prob.update(q=q_host[i & 1])                 # H2D 8n bytes
        res = prob.solve()                           # persistent kernel + D2H of x, y (+ certificate slots)
        wall_ms += 1e3 * (time.perf_counter() - t0)
        st = prob.stats()
        dev_ms += st["kernel_ms"]; iters += st["admm_iters"]; cg += st["cg_iters"]; bytes_algo += st["algo_bytes"]
        phases = [a + b for a, b in zip(phases, st["phase_ms"])]
        launches += 1 + 3 + 2                        # ADMM kernel; q-update (scale + 2 norm kernels); store x / y
        status = res.info.status
    clocks = sampler.stop()
    peak, peak_src = measured_peak()
    value = iters / (dev_ms * 1e-3)
    achieved = bytes_algo / (dev_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(), "admm_iters_per_step": iters // args.steps,
                   "pcg_iters_per_admm_iter": cg / max(iters, 1), "status": status, "l2": "inputs larger than L2 "
                   "(working set ~0.45 GB vs 126 MB L2); no flush between steps",
                   "phase_ms_per_step": dict(zip(["R", "C1_Ap", "C2_Kp", "C3_axpy_dots", "C4_p", "Z", "x6", "x7"],
                                                 [round(v / args.steps, 3) for v in phases])),
                   "grid": f"{st['grid']}x{st['block']}", "setup_s": t_setup, "problem_gen_s": t_gen,
                   "settings": {k: v for k, v in SOLVE_KW.items()}},
        "roofline": {"bound": "hbm", "kernel": "b200qp::admm_kernel (persistent, one launch per solve)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "frac_of_8TBs": achieved / 8000.0, "peak_source": peak_src, "traffic": NCU_TRAFFIC_BYTES,
                     "traffic_source": NCU_TRAFFIC_SOURCE,
                     "algo_bytes_per_launch": bytes_algo / args.steps, "ms_per_launch": dev_ms / args.steps},
        "e2e": {"value": iters / (wall_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * n,
                "d2h_bytes_per_step": 16 * (n + m) + 128, "ms_per_step": wall_ms / args.steps,
                "api": "osqp.OSQP.update(q=host) + solve() -> res.x, res.y on host"},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    # the secondary lines must never cost the headline line: a failure is recorded, not raised
    def guarded(key, fn):
        try:
            line[key] = fn()
        except Exception as exc:      # noqa: BLE001
            line[key] = {"error": f"{type(exc).__name__}: {exc}"}
        return "error" not in line[key]

    if not args.no_spmv:
        guarded("spmv", lambda: spmv_line(A, peak))
    if not args.no_structured:
        if guarded("structured", lambda: structured_line(args, peak)):
            line["gpu_launches"] += 5
    if not args.no_batch:
        from b200qp import batchbench
        if guarded("batch", lambda: batchbench.batch_line(args, with_cpu=not args.no_cpu_baseline)):
            line["gpu_launches"] += args.steps * line["batch"]["gpu_launches_per_step"]
    if not args.no_cpu_baseline:
        if not guarded("cpu_baseline", lambda: cpu_baseline_sample(P, q, A, l, u)):
            line["cpu_baseline"].update({"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed"})
    else:
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "skipped (--no-cpu-baseline)"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-spmv", action="store_true")
    ap.add_argument("--no-batch", action="store_true", help="skip the config-5 (batched MPC QPs) line")
    ap.add_argument("--no-structured", action="store_true", help="skip the banded-sparsity counterpart of config 4")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
