import sys, numpy as np
sys.path.insert(0, '.')
from mpbp_b200 import _lib
L = _lib.lib()
rng = np.random.default_rng(0)
p, n, d, decay, batch = 200, 400, 20, 0.85, 148
c = min(p, n)
Uq, _ = np.linalg.qr(rng.standard_normal((p, c))); Vq, _ = np.linalg.qr(rng.standard_normal((n, c)))
M = (Uq * decay ** np.arange(c)) @ Vq.T
Mall = np.ascontiguousarray(np.broadcast_to(M.T, (batch, n, p)))
U = np.zeros((batch, d, p)); S = np.zeros((batch, d)); st = np.zeros(5); ms = np.zeros(1)
_lib.check(L.mpbp_test_svd(Mall.ctypes.data_as(_lib.c_dp), batch, p, n, d, U.ctypes.data_as(_lib.c_dp), S.ctypes.data_as(_lib.c_dp), st.ctypes.data_as(_lib.c_dp), ms.ctypes.data_as(_lib.c_dp)))
print(ms[0], st)
