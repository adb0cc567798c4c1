import sys, numpy as np
sys.path.insert(0, '.')
from mpbp_b200 import _lib
L = _lib.lib()
np.set_printoptions(linewidth=200, precision=4)
for (m, n, H) in [(3, 5, 32), (3, 12, 32), (9, 9, 32)]:
    rng = np.random.default_rng(m * 1000 + n)
    A = rng.standard_normal((1, m, n)); R = np.zeros((1, n, n)); ms = np.zeros(1)
    _lib.check(L.mpbp_test_qr_ft(A.ctypes.data_as(_lib.c_dp), 1, m, n, H, R.ctypes.data_as(_lib.c_dp), ms.ctypes.data_as(_lib.c_dp)))
    print(m, n); print(R[0]); print(np.linalg.qr(A[0], mode='r'))
    G = A[0].T @ A[0]; k = min(m, n); print("gram err", np.max(np.abs(R[0][:k].T @ R[0][:k] - G)), "full gram err", np.max(np.abs(R[0].T @ R[0] - G)))
