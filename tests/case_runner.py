"""Shared driver: replay a golden case (tests/golden/cases/*.npz) on any solver that offers the
reference call surface (setup / update / warm_start / solve).  Used with the CPU oracle in the
`not gpu` tests and with the CUDA engine (through osqp.OSQP -> pybind11 -> C ABI) in the gpu tests."""
import glob
import json
import os
import sys

import numpy as np
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = os.path.join(ROOT, "tests", "golden", "cases")
PKG = os.path.join(ROOT, "osqp-python_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def case_names():
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(CASES, "*.npz")))


def load_case(name):
    d = dict(np.load(os.path.join(CASES, name + ".npz"), allow_pickle=False))
    n, m = int(d["n"]), int(d["m"])
    c = {"name": name, "n": n, "m": m}
    c["P"] = sparse.csc_matrix((d["P_data"], d["P_indices"], d["P_indptr"]), shape=(n, n))
    c["A"] = sparse.csc_matrix((d["A_data"], d["A_indices"], d["A_indptr"]), shape=(m, n))
    c["q"], c["l"], c["u"] = d["q"], d["l"], d["u"]
    c["settings"] = json.loads(str(d["settings"]))
    c["ops"] = json.loads(str(d["ops"]))
    c["expect"] = {k: v for k, v in d.items() if k.startswith(("ref_", "purepy_"))}
    return c


def run_case_oracle(case, linsys="direct", **override):
    from oracle.osqp_oracle import OracleSolver
    st = dict(case["settings"])
    st.update(override)
    s = OracleSolver(case["P"], case["q"], case["A"], case["l"], case["u"], linsys=linsys, **st)
    res = None
    for kind, kw in case["ops"]:
        if kind == "solve":
            res = s.solve()
        elif kind == "warm_start":
            s.warm_start(x=np.array(kw["x"], float), y=np.array(kw["y"], float))
        elif kind == "update":
            kw = {k: np.array(v) for k, v in kw.items()}
            s.update(P_pattern=case["P"], A_pattern=case["A"], **kw)
    return res


def run_case_engine(case, **override):
    """Through the reference-facing facade: osqp.OSQP(algebra='cuda')."""
    import osqp
    st = dict(case["settings"])
    st.update(override)
    st.setdefault("verbose", False)
    prob = osqp.OSQP(algebra="cuda")
    prob.setup(P=case["P"], q=case["q"], A=case["A"], l=case["l"], u=case["u"], **st)
    res = None
    for kind, kw in case["ops"]:
        if kind == "solve":
            res = prob.solve()
        elif kind == "warm_start":
            prob.warm_start(x=np.array(kw["x"], float), y=np.array(kw["y"], float))
        elif kind == "update":
            kw = {k: (np.array(v, dtype=np.int32) if k.endswith("_idx") else np.array(v, float))
                  for k, v in kw.items()}
            prob.update(**kw)
    return res, prob


def status_name(res):
    s = res.info.status
    return s.decode() if isinstance(s, bytes) else str(s)
