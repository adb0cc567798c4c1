import numpy as np, sys
sys.path.insert(0,'/root/repo')
from mpbp_b200 import _lib
def run(M):
    p,c=M.shape
    A=np.ascontiguousarray(M.T.copy()[None])  # col-major p x c == row-major c x p
    sig=np.zeros((1,c)); order=np.zeros((1,c),dtype=np.int32)
    _lib.check(_lib.lib().mpbp_test_jacobi(A.ctypes.data_as(_lib.c_dp),1,p,c,sig.ctypes.data_as(_lib.c_dp),order.ctypes.data_as(_lib.c_i32p)))
    return sig[0], A[0].T
rng=np.random.default_rng(0)
u,v=rng.standard_normal(4),rng.standard_normal(4)
for M in [np.stack([u,v,2*u,-3*v],1), np.stack([u,u,u,u],1), np.stack([u,v,u+v,u-v],1), np.stack([u,0*u,v,0*v],1),
          np.stack([u,v,2*u,-3*v,u+v,v],1)[:, :6]]:
    s,A=run(M)
    print(M.shape, s, np.linalg.svd(M,compute_uv=False))
