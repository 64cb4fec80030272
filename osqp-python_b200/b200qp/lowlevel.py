"""ctypes view of lib/libb200qp.so (include/b200qp.h) -- what a maintainer's FFI stub binds.
Used by the tests (symbol check, SpMV parity through the raw C ABI) and by bench.py's SpMV line.
Raises if the library is missing: there is no fallback."""
import ctypes as C
import os

import numpy as np

_LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lib", "libb200qp.so")
_lib = None

f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class Settings(C.Structure):
    """b200qp_settings (29 fields, same order as include/b200qp.h)."""
    _fields_ = [("device", C.c_int), ("linsys_solver", C.c_int), ("verbose", C.c_int), ("warm_starting", C.c_int),
                ("scaling", C.c_int), ("polishing", C.c_int), ("rho", C.c_double), ("rho_is_vec", C.c_int),
                ("sigma", C.c_double), ("alpha", C.c_double), ("cg_max_iter", C.c_int), ("cg_tol_reduction", C.c_int),
                ("cg_tol_fraction", C.c_double), ("cg_precond", C.c_int), ("adaptive_rho", C.c_int),
                ("adaptive_rho_interval", C.c_int), ("adaptive_rho_fraction", C.c_double),
                ("adaptive_rho_tolerance", C.c_double), ("max_iter", C.c_int), ("eps_abs", C.c_double),
                ("eps_rel", C.c_double), ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
                ("scaled_termination", C.c_int), ("check_termination", C.c_int), ("check_dualgap", C.c_int),
                ("time_limit", C.c_double), ("delta", C.c_double), ("polish_refine_iter", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} not built (python __graft_entry__.py); the engine has no fallback path")
        L = C.CDLL(_LIB_PATH)
        L.b200qp_version.restype = C.c_char_p
        L.b200qp_error_message.restype = C.c_char_p
        L.b200qp_spmv_variant_name.restype = C.c_char_p
        L.b200qp_spmv_csr_host.argtypes = [C.c_int, C.c_int, i32p, i32p, f64p, f64p, f64p, C.c_int]
        L.b200qp_spmv_plan_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, i32p, i32p, f64p, C.c_int]
        L.b200qp_spmv_plan_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                           C.POINTER(C.c_double)]
        L.b200qp_spmv_plan_destroy.argtypes = [C.c_void_p]
        L.b200qp_set_default_settings.argtypes = [C.POINTER(Settings)]
        _lib = L
    return _lib


def default_settings():
    s = Settings()
    lib().b200qp_set_default_settings(C.byref(s))
    return s


def spmv_host(M, x, variant=0):
    """y = M x through b200qp_spmv_csr_host (M: scipy CSR)."""
    M = M.tocsr()
    y = np.zeros(M.shape[0])
    idx = np.ascontiguousarray(M.indices, np.int32) if M.nnz else np.zeros(1, np.int32)
    val = np.ascontiguousarray(M.data, np.float64) if M.nnz else np.zeros(1)
    rc = lib().b200qp_spmv_csr_host(M.shape[0], M.shape[1], np.ascontiguousarray(M.indptr, np.int32), idx, val,
                                    np.ascontiguousarray(x, np.float64), y if len(y) else np.zeros(1), variant)
    if rc:
        raise RuntimeError(f"b200qp_spmv_csr_host failed: {lib().b200qp_error_message(rc).decode()}")
    return y


class SpmvPlan:
    """Device-resident CSR matrix + timed launches (b200qp_spmv_plan_*)."""

    def __init__(self, M, device=0):
        M = M.tocsr()
        self.shape, self.nnz = M.shape, M.nnz
        self._h = C.c_void_p()
        rc = lib().b200qp_spmv_plan_create(C.byref(self._h), M.shape[0], M.shape[1],
                                           np.ascontiguousarray(M.indptr, np.int32),
                                           np.ascontiguousarray(M.indices, np.int32),
                                           np.ascontiguousarray(M.data, np.float64), device)
        if rc:
            raise RuntimeError(f"b200qp_spmv_plan_create failed: {lib().b200qp_error_message(rc).decode()}")

    def run(self, x=None, variant=0, reps=1, flush_l2=False, want_y=True):
        y = np.zeros(self.shape[0]) if want_y else None
        ms = C.c_double(0.0)
        xp = np.ascontiguousarray(x, np.float64).ctypes.data if x is not None else None
        rc = lib().b200qp_spmv_plan_run(self._h, xp, y.ctypes.data if want_y else None, variant, reps,
                                        1 if flush_l2 else 0, C.byref(ms))
        if rc:
            raise RuntimeError(f"b200qp_spmv_plan_run failed: {lib().b200qp_error_message(rc).decode()}")
        return y, ms.value

    def algo_bytes(self):
        r, c = self.shape
        return 12.0 * self.nnz + 4.0 * (r + 1) + 8.0 * r + 8.0 * c

    def close(self):
        if self._h:
            lib().b200qp_spmv_plan_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close


def variants():
    L = lib()
    return [L.b200qp_spmv_variant_name(i).decode() for i in range(L.b200qp_spmv_num_variants())]
