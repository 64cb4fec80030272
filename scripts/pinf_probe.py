import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/osqp-python_b200')
import numpy as np, osqp
from scipy import sparse
np.random.seed(4)
n, m = 50, 500
Pt = np.random.rand(n, n); P = sparse.triu(Pt.T.dot(Pt), format='csc'); q = np.random.rand(n)
A = sparse.random(m, n).tolil(); u = 3 + np.random.randn(m); l = -3 + np.random.randn(m)
A[int(n/2), :] = A[int(n/2)+1, :]; l[int(n/2)] = u[int(n/2)+1] + 10*np.random.rand(); u[int(n/2)] = l[int(n/2)] + 0.5
A = A.tocsc()
ref = np.load('/root/repo/tests/golden/ref_solutions/test_primal_infeasibility.npz')['normalized_prim_inf_cert_correct']
for kw in [dict(), dict(check_termination=1), dict(cg_tol_fraction=1e-3), dict(cg_tol_fraction=1e-6, cg_max_iter=500)]:
    prob = osqp.OSQP(); prob.setup(P, q, A, l, u, verbose=False, eps_abs=1e-5, eps_rel=1e-5, eps_dual_inf=1e-20, max_iter=2500, polishing=False, **kw)
    r = prob.solve()
    c = r.prim_inf_cert / np.linalg.norm(r.prim_inf_cert)
    d = np.abs(c - ref)
    print(kw, r.info.status, r.info.iter, "max diff", d.max(), "nnz ref", np.count_nonzero(ref), "nnz ours", np.count_nonzero(c), "worst idx", int(d.argmax()), ref[d.argmax()], c[d.argmax()])
    big = np.argsort(-d)[:6]
    print("   ", [(int(i), float(ref[i]), float(c[i])) for i in big])
