"""micro-benchmark of the flat-tree DMMA QR kernel: TFLOP/s at the shapes of the hot path"""
import sys, numpy as np
sys.path.insert(0, '.')
from mpbp_b200 import _lib
L = _lib.lib()
peak = np.zeros(1); _lib.check(L.mpbp_measure_fp64_peak(0, peak.ctypes.data_as(_lib.c_dp)))
print("fp64 dmma peak TF/s", peak[0])
rng = np.random.default_rng(0)
for (m, n, H, batch) in [(1600, 400, 32, 296), (4000, 400, 32, 296), (1600, 400, 64, 296), (4000, 400, 64, 296), (1600, 400, 64, 148), (1600, 400, 32, 148), (8000, 400, 64, 148), (400, 100, 32, 1184), (400, 100, 64, 1184)]:
    A = rng.standard_normal((batch, m, n)); R = np.zeros((batch, n, n)); ms = np.zeros(1)
    _lib.check(L.mpbp_test_qr_ft(A.ctypes.data_as(_lib.c_dp), batch, m, n, H, R.ctypes.data_as(_lib.c_dp), ms.ctypes.data_as(_lib.c_dp)))
    fl = batch * (2.0 * m * n * n - 2.0 / 3.0 * n ** 3)
    G = A[0].T @ A[0]; err = np.max(np.abs(R[0].T @ R[0] - G)) / np.abs(G).max()
    print(f"m={m} n={n} H={H} batch={batch}: {ms[0]:.2f} ms  {fl / ms[0] / 1e9:.2f} TF/s  ({100 * fl / ms[0] / 1e9 / peak[0]:.1f}% of peak) err={err:.1e}")
