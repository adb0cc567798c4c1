"""A/B of builds of the flat-tree QR kernel: python tools/qr_variants.py lib1.so lib2.so ...  (one process per lib)"""
import subprocess, sys
if len(sys.argv) > 2:
    for p in sys.argv[1:]:
        subprocess.run([sys.executable, __file__, p], check=False)
    sys.exit(0)
import numpy as np
sys.path.insert(0, '.')
from mpbp_b200 import _lib
if len(sys.argv) == 2:
    _lib.LIB_PATH = sys.argv[1]
L = _lib.lib()
rng = np.random.default_rng(0)
out = []
for (m, n, H, batch) in [(1600, 400, 64, 148), (4000, 400, 64, 148), (1600, 400, 64, 1), (3200, 400, 64, 296)]:
    A1 = rng.standard_normal((m, n))
    A = np.ascontiguousarray(np.broadcast_to(A1, (batch, m, n)))
    R = np.zeros((batch, n, n)); ms = np.zeros(1)
    _lib.check(L.mpbp_test_qr_ft(A.ctypes.data_as(_lib.c_dp), batch, m, n, H, R.ctypes.data_as(_lib.c_dp), ms.ctypes.data_as(_lib.c_dp)))
    fl = batch * (2.0 * m * n * n - 2.0 / 3.0 * n ** 3)
    G = A1.T @ A1; err = np.max(np.abs(R[-1].T @ R[-1] - G)) / np.abs(G).max()
    out.append(f"{m}x{n} b{batch}: {ms[0]:.2f} ms {fl / ms[0] / 1e9:.2f} TF/s err={err:.0e}")
print(sys.argv[1] if len(sys.argv) == 2 else "default", "|", " | ".join(out))
