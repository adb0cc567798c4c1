"""one launch set of the product QR device code (H = 64, 148 matrices 1600 x 400) with a given build of the library:
python tools/qr_one_lib.py path/to/lib.so   (used under ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum)"""
import sys, numpy as np
sys.path.insert(0, '.')
from mpbp_b200 import _lib
if len(sys.argv) == 2:
    _lib.LIB_PATH = sys.argv[1]
L = _lib.lib()
m, n, H, batch = 1600, 400, 64, 148
rng = np.random.default_rng(0)
A = rng.standard_normal((batch, m, n)); R = np.zeros((batch, n, n)); ms = np.zeros(1)
_lib.check(L.mpbp_test_qr_ft(A.ctypes.data_as(_lib.c_dp), batch, m, n, H, R.ctypes.data_as(_lib.c_dp), ms.ctypes.data_as(_lib.c_dp)))
print("ms", ms[0], "TF/s", batch * (2.0 * m * n * n - 2.0 / 3.0 * n ** 3) / ms[0] / 1e9)
