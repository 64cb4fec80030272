"""The reference's own test files, unmodified, through the reference's own Python facade (tests/ref/README.md).

Two bindings under that facade:
  optionB  this repository's pybind11 module  osqp-python_b200/osqp_cuda*.so      (csrc/bindings.cpp)
  optionA  the reference's UNMODIFIED src/bindings.cpp.in compiled over include/osqp_shim/*.h
Both resolve `algebra='cuda'` (interface.py:20-24) to libb200qp.so.  The suite's conftest then yields its
('cuda', 'indirect', 1e-2, 1e-3, 2) row (conftest.py:26-29)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORK = os.path.join(ROOT, "tests", "ref", "_work")
PKG = os.path.join(WORK, "pkg")
BINDING_DIR = {"optionB": os.path.join(ROOT, "osqp-python_b200"), "optionA": os.path.join(WORK, "optionA")}

needs_work = pytest.mark.skipif(not os.path.isdir(os.path.join(PKG, "osqp", "tests")),
                                reason="tests/ref/_work not assembled (run __graft_entry__.build() where /root/reference exists)")


def _env(which):
    env = dict(os.environ)
    # the reference facade first, then the directory of the extension module; nothing of osqp-python_b200/osqp/ is used:
    # `import osqp` resolves to _work/pkg/osqp because it comes first on the path
    env["PYTHONPATH"] = os.pathsep.join([PKG, BINDING_DIR[which]])
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    return env


def _run(which, code):
    return subprocess.run([sys.executable, "-c", code], env=_env(which), capture_output=True, text=True, timeout=600)


@needs_work
@pytest.mark.parametrize("which", ["optionB", "optionA"])
def test_reference_facade_resolves_to_the_engine(which):
    """No GPU needed: the unmodified facade imports, finds exactly the 'cuda' algebra, builds its enums from the module
    and reaches the engine's setup (which fails loudly without a device: OSQP_ALGEBRA_LOAD_ERROR, never a CPU path)."""
    code = r'''
import numpy as np, scipy.sparse as sp, os, osqp, osqp_cuda
assert os.path.realpath(osqp.__file__).startswith(os.path.realpath(%r)), osqp.__file__
assert os.path.realpath(osqp_cuda.__file__).startswith(os.path.realpath(%r)), osqp_cuda.__file__
assert osqp.algebras_available() == ['cuda'] and osqp.default_algebra() == 'cuda'
assert osqp.SolverStatus.OSQP_SOLVED == 1 and osqp.SolverError.OSQP_ALGEBRA_LOAD_ERROR == 7
caps = osqp_cuda.osqp_capabilities()
assert caps & osqp_cuda.osqp_capabilities_type.OSQP_CAPABILITY_INDIRECT_SOLVER.value
assert caps & osqp_cuda.osqp_capabilities_type.OSQP_CAPABILITY_UPDATE_MATRICES.value
m = osqp.OSQP()
P = sp.csc_matrix([[4.0, 1], [1, 2]]); q = np.array([1.0, 1]); A = sp.csc_matrix([[1.0, 1], [1, 0], [0, 1]])
l = np.array([1.0, 0, 0]); u = np.array([1.0, 0.7, 0.7])
try:
    m.setup(P, q, A, l, u, verbose=False)
    r = m.solve()
    assert r.info.status == 'solved' and abs(r.info.obj_val - 1.88) < 1e-2, (r.info.status, r.info.obj_val)
    print('SOLVED-ON-GPU')
except osqp.OSQPException as e:
    assert e == osqp.SolverError.OSQP_ALGEBRA_LOAD_ERROR, e.args
    print('NO-DEVICE')
''' % (PKG, BINDING_DIR[which])
    r = _run(which, code)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SOLVED-ON-GPU" in r.stdout or "NO-DEVICE" in r.stdout


# Tests of the suite whose `cuda` / indirect branch pins a known DEFECT of the reference's iterative back ends
# ("# pytest-todo" in the reference); a solver that gets the right answer cannot pass them:
#   basic_test.py:94-99        the feasible problem of test_update_bounds must come back PRIMAL_INFEASIBLE
#   feasibility_test.py:51-56  an indirect solver must fail to converge (MAX_ITER_REACHED after 5000 iterations)
# This engine returns the golden optimum in both cases, which tests/test_gpu_parity.py::test_engine_vs_reference_golden
# [test_update_bounds] / [test_feasibility_problem] check at the reference's strict (builtin) tolerances.
KNOWN_DEFECT_PINS = ["osqp/tests/basic_test.py::test_update_bounds",
                     "osqp/tests/feasibility_test.py::test_feasibility_problem"]
# primal_infeasibility_test.py:58 compares a certificate with np.allclose DEFAULTS (rtol 1e-5, atol 1e-8) against the
# iterate a *direct* solver stopped at: entries of size 1e-5..1e-6 next to one of size 1.  That pins the whole ADMM
# trajectory, i.e. needs inner solves exact to ~1e-7; with the engine's default inexact PCG (cg_tol_fraction 0.1) those
# entries differ by 2e-5 (the restated test, tests/test_gpu_parity.py, allows 1e-4).  So the suite runs twice:
#   "exact"   B200QP_CG_TOL_FRACTION=1e-7: every test of the suite (minus the defect pins)
#   "default" engine defaults: everything but that one certificate comparison
TRAJECTORY_PIN = "osqp/tests/primal_infeasibility_test.py::test_primal_infeasible_problem"


@needs_work
@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["exact", "default"])
@pytest.mark.parametrize("which", ["optionB", "optionA"])
def test_reference_suite_unmodified(which, mode):
    cmd = [sys.executable, "-m", "pytest", "osqp/tests", "-q", "--no-header", "-p", "no:cacheprovider", "-W", "ignore"]
    env = _env(which)
    for t in KNOWN_DEFECT_PINS + ([TRAJECTORY_PIN] if mode == "default" else []):
        cmd += ["--deselect", t]
    if mode == "exact":
        env["B200QP_CG_TOL_FRACTION"] = "1e-7"
        env["B200QP_CG_MAX_ITER"] = "500"
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900, cwd=PKG)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and " failed" not in r.stdout, tail
    print(f"[refsuite {which} {mode}] " + r.stdout.strip().splitlines()[-1])


EXAMPLES = ["basic_usage.py", "update_vectors.py", "update_matrices.py", "exception_handling.py"]


@needs_work
@pytest.mark.gpu
@pytest.mark.parametrize("script", EXAMPLES)
@pytest.mark.parametrize("which", ["optionB", "optionA"])
def test_reference_examples_run_unmodified(which, script):
    """The reference's example scripts (/root/reference/examples, copied into the work tree at build time) run as they
    are through the reference facade over this engine.  exception_handling.py catches the OSQP_NONCVX_ERROR a *direct*
    solver raises at setup; an indirect engine raises nothing there (non_convex_test.py: the `else` branch), the script
    simply falls through."""
    path = os.path.join(WORK, "examples", script)
    if not os.path.exists(path):
        pytest.skip("examples not in the work tree (rebuild where /root/reference exists)")
    r = subprocess.run([sys.executable, path], env=_env(which), capture_output=True, text=True, timeout=300, cwd=WORK)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    if script != "exception_handling.py":
        assert "solved" in r.stdout, r.stdout[-1500:]
