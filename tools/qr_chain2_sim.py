"""numpy model of the schedule of tools/qr_chain2.cu (which panel is applied to which slab, by whom, in which step):
checks that the look-ahead-depth-2 / streamed-reflector ordering yields R with R^T R = A^T A.  Run: python tools/qr_chain2_sim.py"""
import numpy as np
rng=np.random.default_rng(0)
m,n,H,B=200,40,64,8
A=rng.standard_normal((m,n))
R=np.zeros((n,n))
npan=n//B
def factor(Rjj, a):
    # Householder on [Rjj(8x8 upper); a (H x 8)], reflectors [e_k; v]; returns new Rjj, V (H x 8), taus
    a=a.copy(); Rn=np.zeros((B,B)); V=np.zeros((a.shape[0],B)); taus=np.zeros(B)
    for k in range(B):
        alpha=Rjj[k,k]; sig2=a[:,k]@a[:,k]
        rk=Rjj[k].copy()
        if sig2>0:
            nrm=np.sqrt(alpha*alpha+sig2); beta=-nrm if alpha>=0 else nrm
            u=alpha-beta; tau=(beta-alpha)/beta; sc=1/u
        else:
            beta=alpha; tau=0; sc=0
        v=a[:,k]*sc; a[:,k]=v; V[:,k]=v; taus[k]=tau
        for c in range(k+1,B):
            s=tau*(rk[c]+v@a[:,c]); a[:,c]-=s*v; rk[c]-=s
        Rn[k,k]=beta; Rn[k,k+1:]=rk[k+1:]
    return Rn,V,taus
def T_of(V,taus):
    T=np.zeros((B,B))
    for k in range(B):
        T[k,k]=taus[k]
        if k: T[:k,k]=-taus[k]*T[:k,:k]@(V[:,:k].T@V[:,k])
    return T
def wy_apply(V,T,Rrows,slab):
    W=Rrows+V.T@slab; Wp=T.T@W
    return Rrows-Wp, slab-V@Wp
def stream(V,taus,Rrows,slab):
    Rrows=Rrows.copy(); slab=slab.copy()
    for k in range(B):
        w=taus[k]*(Rrows[k]+V[:,k]@slab); slab-=np.outer(V[:,k],w); Rrows[k]-=w
    return Rrows,slab
for row0 in range(0,m,H):
    blk=np.zeros((H,n)); rows=A[row0:row0+H]; blk[:rows.shape[0]]=rows
    regs={0:blk[:,0:B].copy()}   # chain warp registers: slab index -> data
    Vs={}; Ts={}; taus={}
    for s in range(npan):
        j0=s*B
        # consumer part (i),(ii) happen at the start of step s using panel s-1
        if s+1<npan:
            c0=j0+B
            if s>=1:
                Rr,sl=wy_apply(Vs[s-1],Ts[s-1],R[j0-B:j0,c0:c0+B],blk[:,c0:c0+B]); R[j0-B:j0,c0:c0+B]=Rr; blk[:,c0:c0+B]=sl
            regs[s+1]=blk[:,c0:c0+B].copy()
        # update warps: panel s-1 on slabs >= s+2
        if s>=1:
            for p in range(s+2,npan):
                c0=p*B
                Rr,sl=wy_apply(Vs[s-1],Ts[s-1],R[j0-B:j0,c0:c0+B],blk[:,c0:c0+B]); R[j0-B:j0,c0:c0+B]=Rr; blk[:,c0:c0+B]=sl
        # producer factors panel s from registers
        Rn,V,ta=factor(R[j0:j0+B,j0:j0+B],regs[s]); R[j0:j0+B,j0:j0+B]=Rn; Vs[s]=V; taus[s]=ta; Ts[s]=T_of(V,ta)
        # consumer (iii) streams panel s onto slab s+1
        if s+1<npan:
            c0=j0+B
            Rr,sl=stream(V,ta,R[j0:j0+B,c0:c0+B],regs[s+1]); R[j0:j0+B,c0:c0+B]=Rr; regs[s+1]=sl
G=A.T@A
print("err",np.abs(np.triu(R).T@np.triu(R)-G).max()/np.abs(G).max(), "lower", np.abs(np.tril(R,-1)).max())
