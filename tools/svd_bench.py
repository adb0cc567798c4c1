"""micro-benchmark of the op-truncation SVD core at hot-path shapes (batch = one CTA per SM)"""
import sys, numpy as np
sys.path.insert(0, '.')
from mpbp_b200 import _lib
L = _lib.lib()
rng = np.random.default_rng(0)
for (p, n, d, decay) in [(200, 400, 20, 0.85), (200, 400, 20, 0.93), (80, 400, 20, 0.85), (440, 400, 20, 0.9)]:
    batch, c = 148, min(p, n)
    Uq, _ = np.linalg.qr(rng.standard_normal((p, c))); Vq, _ = np.linalg.qr(rng.standard_normal((n, c)))
    s = decay ** np.arange(c)
    M = (Uq * s) @ Vq.T
    Mall = np.ascontiguousarray(np.broadcast_to(M.T, (batch, n, p)))
    U = np.zeros((batch, d, p)); S = np.zeros((batch, d)); st = np.zeros(5); ms = np.zeros(1)
    _lib.check(L.mpbp_test_svd(Mall.ctypes.data_as(_lib.c_dp), batch, p, n, d, U.ctypes.data_as(_lib.c_dp), S.ctypes.data_as(_lib.c_dp), st.ctypes.data_as(_lib.c_dp), ms.ctypes.data_as(_lib.c_dp)))
    Un, sn, _ = np.linalg.svd(M, full_matrices=False); Ud = U[0].T
    err = np.linalg.norm(Un[:, :d] @ (Un[:, :d].T @ M) - Ud @ (Ud.T @ M)) / sn[0]
    print(f"p={p} n={n} d={d} decay={decay}: {ms[0]:.2f} ms per wave of {batch}; calls={st[0]:.0f} iters/call={st[1]/max(st[0],1):.1f} b={st[2]/max(st[0],1):.0f} maxsweeps/call={st[3]/max(st[0],1):.1f} unconv={st[4]:.0f} proj_err={err:.1e}")
