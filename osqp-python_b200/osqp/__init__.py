"""`osqp` -- the reference's Python surface (src/osqp/__init__.py:1-13) over the B200 engine.

Only the solve() path is provided: OSQP().setup / solve / update / update_settings / warm_start,
constant(), SolverStatus, SolverError, OSQPException.  There is one algebra, 'cuda', served by the
extension module `osqp_cuda` (pybind11 -> include/b200qp.h C ABI -> hand-written sm_100a kernels).
"""
from osqp._version import version as __version__  # noqa: F401
from osqp.interface import (  # noqa: F401
    OSQP,
    OSQPException,
    SolverError,
    SolverStatus,
    algebra_available,
    algebras_available,
    constant,
    default_algebra,
)
from osqp.batch import OSQPBatch  # noqa: F401,E402  (batched front end, BASELINE.json configs[4])
