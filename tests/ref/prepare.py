"""Assemble tests/ref/_work/ (git-ignored; travels to the GPU box with the snapshot).  See tests/ref/README.md.

Needs /root/reference (present in the build container only); a no-op elsewhere when _work already exists."""
import glob
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
WORK = os.path.join(HERE, "_work")
REF = os.environ.get("OSQP_REFERENCE_ROOT", "/root/reference")
FACADE_FILES = ("__init__.py", "interface.py", "builtin.py", "cuda.py", "mkl.py")


def prepare(verbose=True):
    src_pkg = os.path.join(REF, "src", "osqp")
    binding_in = os.path.join(REF, "src", "bindings.cpp.in")
    if not os.path.isdir(src_pkg):
        if verbose:
            print(f"[tests/ref] {REF} not present: keeping the existing {WORK}" if os.path.isdir(WORK)
                  else f"[tests/ref] {REF} not present and no _work tree: the reference-suite tests will skip")
        return os.path.isdir(WORK)
    pkg = os.path.join(WORK, "pkg", "osqp")
    shutil.rmtree(os.path.join(WORK, "pkg"), ignore_errors=True)
    os.makedirs(os.path.join(pkg, "tests", "solutions"))
    for f in FACADE_FILES:                                   # the reference facade: copied at build time, never committed
        shutil.copy(os.path.join(src_pkg, f), os.path.join(pkg, f))
    with open(os.path.join(pkg, "_version.py"), "w") as fh:  # setuptools-scm artefact of the reference build
        fh.write("version = '1.0.0+b200qp'\n")
    open(os.path.join(pkg, "tests", "__init__.py"), "w").close()
    for f in glob.glob(os.path.join(HERE, "suite", "*.py")):
        shutil.copy(f, os.path.join(pkg, "tests"))
    for f in glob.glob(os.path.join(ROOT, "tests", "golden", "ref_solutions", "*.npz")):
        shutil.copy(f, os.path.join(pkg, "tests", "solutions"))
    # the reference's in-scope example scripts, copied at build time like the facade, never committed
    ex = os.path.join(WORK, "examples")
    shutil.rmtree(ex, ignore_errors=True)
    os.makedirs(ex)
    for f in ("basic_usage.py", "update_vectors.py", "update_matrices.py", "exception_handling.py"):
        shutil.copy(os.path.join(REF, "examples", f), os.path.join(ex, f))
    # Option A: the reference's unmodified binding source over the shim headers
    optA = os.path.join(WORK, "optionA")
    os.makedirs(optA, exist_ok=True)
    gen = os.path.join(optA, "bindings_ref.cpp")
    with open(binding_in) as fh:
        text = fh.read().replace("@OSQP_EXT_MODULE_NAME@", "osqp_cuda")
    with open(gen, "w") as fh:
        fh.write(text)
    import pybind11
    libdir = os.path.join(ROOT, "osqp-python_b200", "lib")
    out = os.path.join(optA, "osqp_cuda" + sysconfig.get_config_var("EXT_SUFFIX"))
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
           "-I", os.path.join(ROOT, "include", "osqp_shim"), "-I", pybind11.get_include(),
           "-I", sysconfig.get_paths()["include"], gen, "-o", out, "-L", libdir, "-lb200qp",
           "-Wl,-rpath,$ORIGIN/../../../../osqp-python_b200/lib"]
    if verbose:
        print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    os.remove(gen)                                           # the instantiated reference source is not kept
    return True


if __name__ == "__main__":
    sys.exit(0 if prepare() else 1)
