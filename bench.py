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

import bench_cpu as CPU  # noqa: E402

METRIC = "admm_iters_per_sec"
UNIT = "ADMM iters/s"
N_VARS = int(os.environ.get("B200QP_BENCH_N", 1_000_000))
N_CONS = int(os.environ.get("B200QP_BENCH_M", N_VARS))
NNZ_ROW = 10
SEED = 3
SOLVE_KW = dict(eps_abs=1e-3, eps_rel=1e-3, max_iter=4000, warm_starting=False, polishing=False)
CPU_SAMPLE_ITERS = int(os.environ.get("B200QP_CPU_SAMPLE_ITERS", 10))    # ADMM iterations per CPU sample (both CPU legs)


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the hot kernel on this workload, from the
    `ncu --set full` capture of the committed build: scripts/ncu_traffic.sh writes profiles/r2_traffic.json next to
    the raw csv it was read from (a number taken under the profiler is never a bench value; it only sizes the traffic)."""
    if (N_VARS, N_CONS) != (1_000_000, 1_000_000):
        return None, "not captured for this problem size"
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        return float(t["dram_bytes_per_launch"]), t["source"]
    except Exception:      # noqa: BLE001
        return None, "profiles/r2_traffic.json missing"


def config_block():
    """Identical in both arms (the driver compares it): what is solved, not what was measured."""
    return {"workload": workload_name(), "n": N_VARS, "m": N_CONS, "nnz_A": N_CONS * NNZ_ROW, "seed": SEED,
            "settings": dict(SOLVE_KW), "engine_spec": dict(CPU.ENGINE_SPEC),
            "l2": "inputs larger than L2 (working set ~0.45 GB vs 126 MB L2); no flush between steps"}


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
    (osqp v1.0.0 / QDLDL) is not vendored and osqppurepy cannot travel to the GPU box, so this is the oracle port of
    it (`kind: "port"`, not the upstream C solver); a direct factorisation cannot finish at n = m = 1e6 on random
    sparsity, hence the engine-spec PCG mode.  One step = one bounded sample: the first CPU_SAMPLE_ITERS
    ADMM iterations of the cold-start solve in ONE solve() call (no per-call overhead per iteration), on every host
    thread (torchrun's OMP_NUM_THREADS=1 is overridden)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    threads = CPU.use_all_host_threads()
    P, q, A, l, u = make_problem()
    budget = float(os.environ.get("B200QP_REF_BUDGET_S", 240))
    t_begin = time.perf_counter()
    done_w = 0
    for _ in range(min(args.warmup, 1)):                       # one short warm-up sample (page-in, OpenMP pool)
        CPU.config4_sample(P, q, A, l, u, SOLVE_KW, 2)
        done_w += 1
    iters = cg = steps = 0
    dt = 0.0
    used = threads
    for _ in range(args.steps):
        if steps and time.perf_counter() - t_begin + dt / steps > budget:
            break
        k, c, t, used = CPU.config4_sample(P, q, A, l, u, SOLVE_KW, CPU_SAMPLE_ITERS)
        iters += k; cg += c; dt += t; steps += 1
    val = iters / dt if dt > 0 else 0.0
    sample = (f"{steps} sample(s) of the first {CPU_SAMPLE_ITERS} ADMM iterations of the cold-start solve ({cg / max(iters, 1):.1f} PCG "
              f"iterations each, {dt:.1f} s in total), one solve() call per sample; oracle port, CSR SpMV on {used} thread(s) "
              f"(OpenMP helper when faster than scipy's csr_matvec on this host) + numpy; {CPU.host_cores()} host cores available")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": done_w, "ms_per_step": 1e3 * dt / max(steps, 1), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(),
            "detail": {"step": f"{CPU_SAMPLE_ITERS} ADMM iterations of the CPU port (bounded sample)",
                       "pcg_iters_per_admm_iter": cg / max(iters, 1)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------- our arm
def cpu_baseline_sample(P, q, A, l, u):
    """Same sampling rule as the reference arm: the first CPU_SAMPLE_ITERS ADMM iterations, one solve() call."""
    k, cg, dt, used = CPU.config4_sample(P, q, A, l, u, SOLVE_KW, CPU_SAMPLE_ITERS)
    return {"value": k / dt, "unit": UNIT, "cores": used, "kind": "port",
            "sample": f"first {k} ADMM iterations ({cg} PCG iterations, {dt:.1f} s) of the same cold-start solve, one solve() call; "
                      f"oracle port, CSR SpMV on {used} thread(s) (OpenMP helper when faster than scipy on this host) + numpy; "
                      f"{CPU.host_cores()} host cores available",
            "pcg_iters_per_admm_iter": cg / max(k, 1)}


def spmv_line(A, peak):
    """Standalone per-iteration SpMV (y = A x) achieved GB/s, every variant; "flushed" = a buffer larger than L2 is READ
    between launches and x is brought back (inside the solver the gathered vector was just written, the matrix is not in
    L2); "tile-tma" is the scheme of the persistent kernel's phases."""
    from b200qp import lowlevel
    M = A.tocsr()
    plan = lowlevel.SpmvPlan(M)
    x = np.random.default_rng(7).standard_normal(M.shape[1])
    out = {}
    for v, name in enumerate(lowlevel.variants()):
        try:
            plan.run(x, variant=v, reps=3, want_y=False)
        except RuntimeError:          # the tile variant refuses matrices it cannot hold (rows > 255 entries)
            continue
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
           "ring": f"{st['ring_stages']} stages x {st['ring_stage_bytes']} B", "ring_autotune_ms": list(st.get("autotune_ms", (0, 0))),
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
    dev_ms, wall_ms, iters, cg, bytes_algo, bytes_engine, launches = 0.0, 0.0, 0, 0, 0.0, 0.0, 0
    status = None
    phases = [0.0] * 8
    for i in range(args.steps):
        t0 = time.perf_counter()
        prob.update(q=q_host[i & 1])                 # H2D 8n bytes
        res = prob.solve()                           # persistent kernel + D2H of the 200-byte result block, x and y
                                                     # (certificate arrays exist only for infeasible statuses: NaN-filled on the host)
        wall_ms += 1e3 * (time.perf_counter() - t0)
        st = prob.stats()
        dev_ms += st["kernel_ms"]; iters += st["admm_iters"]; cg += st["cg_iters"]; bytes_algo += st["algo_bytes"]
        bytes_engine += st["algo_bytes_engine"]
        phases = [a + b for a, b in zip(phases, st["phase_ms"])]
        launches += 1 + 3 + 2                        # ADMM kernel; q-update (scale + 2 norm kernels); store x / y
        status = res.info.status
    clocks = sampler.stop()
    peak, peak_src = measured_peak()
    value = iters / (dev_ms * 1e-3)
    achieved = bytes_algo / (dev_ms * 1e-3) / 1e9
    achieved_engine = bytes_engine / (dev_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic()
    tile = st["kernel_kind"] == 1
    names = ["R", "B_Au", "C_Ku", "V_axpy_dots", "x4", "Z", "x6", "x7"] if tile else ["R", "C1_Ap", "C2_Kp", "C3_axpy_dots", "C4_p", "Z", "x6", "x7"]
    kernel = ("b200qp::admm_kernel_csl (persistent; compacted-slice tiles streamed by cp.async.bulk into an mbarrier ring, "
              "one launch per solve)") if tile else "b200qp::admm_kernel (persistent CSR lanes-per-row kernel, one launch per solve)"
    # the per-iteration SpMV the metric names: in-kernel phase time per PCG iteration (CTA 0's clock, barrier wait included)
    # against the algorithmic bytes of one product
    def phase_roof(ms_total, nbytes):
        us = 1e3 * ms_total / max(cg, 1)
        gbs = nbytes / (us * 1e-6) / 1e9 if us > 0 else 0.0
        return {"us_per_product": us, "algo_bytes": nbytes, "achieved": gbs, "unit": "GB/s", "frac": gbs / peak,
                "frac_of_8TBs": gbs / 8000.0}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_block(),
        "detail": {"admm_iters_per_step": iters // args.steps, "pcg_iters_per_admm_iter": cg / max(iters, 1), "status": status,
                   "phase_ms_per_step": dict(zip(names, [round(v / args.steps, 3) for v in phases])),
                   "grid": f"{st['grid']}x{st['block']}", "ring": f"{st['ring_stages']} stages x {st['ring_stage_bytes']} B" if tile else None,
                   "ring_autotune_ms": list(st.get("autotune_ms", (0, 0))),
                   "setup_s": t_setup, "problem_gen_s": t_gen},
        "roofline": {"bound": "hbm", "kernel": kernel,
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "frac_of_8TBs": achieved / 8000.0, "peak_source": peak_src,
                     "byte_model": "SURVEY.md 8(d): per PCG iteration B(P)+B(A)+B(A')+64n, per ADMM iteration B(A')+B(A)+8(5n+10m), "
                                   "per check B(P)+B(A'); B(M) = 12 nnz + 4(rows+1) + 8 rows + 8 cols",
                     "algo_bytes_per_launch": bytes_algo / args.steps, "ms_per_launch": dev_ms / args.steps,
                     "engine_model": {"achieved": achieved_engine, "frac": achieved_engine / peak,
                                      "algo_bytes_per_launch": bytes_engine / args.steps,
                                      "what": "the kernel's own storage and vector passes (tile blobs, V phase 7 reads + 5 writes)"},
                     "traffic": traffic, "traffic_source": traffic_src,
                     "per_iteration_spmv": {"A_u": phase_roof(phases[1], st["spmv_bytes_A"]),
                                            "K_u": phase_roof(phases[2], st["spmv_bytes_K"]),
                                            "note": "phase time on CTA 0 incl. the grid barrier, per PCG iteration; K u = P u + sigma u + A' t"}},
        "e2e": {"value": iters / (wall_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * n,
                "d2h_bytes_per_step": 8 * (n + m) + 200, "ms_per_step": wall_ms / args.steps,
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
        cpu_fn = None if args.no_cpu_baseline else CPU.mpc_batch_sample
        if guarded("batch", lambda: batchbench.batch_line(args, cpu_sample=cpu_fn)):
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
