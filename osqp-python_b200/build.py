"""Build the engine in-tree (the built .so files travel to the GPU box with the repo snapshot).

  lib/libb200qp.so                      nvcc, sm_100a only, C ABI of include/b200qp.h
  osqp_cuda.cpython-312-*.so            g++ + pybind11, links lib/libb200qp.so via $ORIGIN rpath

    python osqp-python_b200/build.py [--force]
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libb200qp.so")
EXT = os.path.join(HERE, "osqp_cuda" + sysconfig.get_config_var("EXT_SUFFIX"))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false"]


def _digest(sources):
    import hashlib
    h = hashlib.sha256()
    for s in sorted(sources):
        h.update(os.path.basename(s).encode())
        h.update(open(s, "rb").read())
    return h.hexdigest()


def _stale(target, sources):
    """Content-hash stamp (mtimes do not survive the snapshot copy to the GPU box)."""
    stamp = target + ".stamp"
    if not os.path.exists(target) or not os.path.exists(stamp):
        return True
    return open(stamp).read().strip() != _digest(sources)


def _stamp(target, sources):
    with open(target + ".stamp", "w") as f:
        f.write(_digest(sources))


def _deps(unit):
    """The unit plus every local header it (transitively) includes."""
    import re
    seen, todo = [], [unit]
    while todo:
        f = os.path.normpath(todo.pop())
        if f in seen or not os.path.exists(f):
            continue
        seen.append(f)
        for inc in re.findall(r'#include\s+"([^"]+)"', open(f).read()):
            todo.append(os.path.join(os.path.dirname(f), inc))
    return seen


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build(force=False, verbose_ptxas=False):
    os.makedirs(LIBDIR, exist_ok=True)
    cu_sources = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    cu_sources.append(os.path.join(ROOT, "include", "b200qp.h"))
    units = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    objs = [os.path.join(LIBDIR, os.path.basename(u)[:-3] + ".o") for u in units]
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math") and f != "-shared"]
    if verbose_ptxas:
        flags += ["-Xptxas", "-v"]
    # one object per translation unit (rebuilt when the unit or a header it includes changed), compiled
    # side by side, then one link
    procs = []
    for u, o in zip(units, objs):
        deps = _deps(u)
        if force or verbose_ptxas or _stale(o, deps):
            cmd = [NVCC] + flags + ["-c", "-o", o, u]
            print("+", " ".join(cmd), flush=True)
            procs.append((cmd, subprocess.Popen(cmd), o, deps))
    for cmd, pr, o, deps in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
        _stamp(o, deps)
    if procs or force or _stale(LIB, cu_sources):
        _run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs)
        _stamp(LIB, cu_sources)
    # the oracle's C helper (test infrastructure: OpenMP CSR SpMV for the CPU baseline); building the checker is
    # not using it
    osrc = os.path.join(ROOT, "oracle", "spmv_omp.c")
    olib = os.path.join(ROOT, "oracle", "_build", "liboracle_spmv.so")
    if os.path.exists(osrc) and (force or _stale(olib, [osrc])):
        os.makedirs(os.path.dirname(olib), exist_ok=True)
        _run(["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-o", olib, osrc])
        _stamp(olib, [osrc])
    bind_src = os.path.join(CSRC, "bindings.cpp")
    ext_sources = [bind_src, os.path.join(ROOT, "include", "b200qp.h")]
    if force or _stale(EXT, ext_sources):
        import pybind11
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
              "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
              bind_src, "-o", EXT, "-L", LIBDIR, "-lb200qp", "-Wl,-rpath,$ORIGIN/lib"])
        _stamp(EXT, ext_sources)
    return LIB, EXT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose_ptxas="-v" in sys.argv)
