"""numpy feasibility check of the column-pair reduction for the QR panel chain: the dots of
column k+1 after reflector k are derived from the reductions already made for column k (one 16-value tree per two
columns) with a recompute guard against cancellation; prints max |R^T R - G| / max|G| and the number of guarded
recomputations for the standard and the paired chain on random, near-dependent, rank-deficient, graded and tiny blocks."""
import numpy as np
rng=np.random.default_rng(1)
def panel(Rjj,a,pair,guard=1e-3):
    """Householder on [Rjj; a] with reflectors [e_k; v]. pair=True: dots for odd columns derived from the even column's reduction."""
    a=a.copy(); B=8; Rn=np.zeros((B,B)); nrecomp=0
    red_next=None
    for k in range(B):
        rk=Rjj[k].copy()
        if pair and (k%2==1) and red_next is not None:
            red=red_next; 
        else:
            red=np.array([a[:,k]@a[:,c] for c in range(B)])   # one reduction tree: dots of column k with all
            if pair and k%2==0 and k+1<B:
                red2=np.array([a[:,k+1]@a[:,c] for c in range(B)])  # same tree, 16 values
        alpha=rk[k]; sig2=red[k]
        if sig2>0:
            nrm=np.sqrt(alpha*alpha+sig2); beta=-nrm if alpha>=0 else nrm; u=alpha-beta; tau=(beta-alpha)/beta; sc=1/u
        else: beta=alpha;tau=0;sc=0
        v=a[:,k]*sc
        s=np.zeros(B)
        for c in range(k+1,B): s[c]=tau*(rk[c]+sc*red[c])
        if pair and k%2==0 and k+1<B:
            # derive dots of updated column k+1 with updated columns c>k (and with v's of earlier cols not needed except T)
            vv=sc*sc*red[k]; va=sc*red   # v.v and v.a_c (old a_c)
            d=np.zeros(B)
            for c in range(B):
                if c>k:
                    d[c]=red2[c]-s[c]*va[k+1]-s[k+1]*va[c]+s[k+1]*s[c]*vv
                elif c==k:
                    d[c]=sc*(red2[k])-s[k+1]*vv      # v_k . a'_{k+1}  (for T); not used for R
                else:
                    d[c]=red2[c]-s[k+1]*0  # v_c.a'_{k+1}: v_c orthogonal relations ignored (T only)
            # guard: cancellation in the norm
            if d[k+1] < guard*red2[k+1]:
                red_next=None; nrecomp+=1
            else: red_next=d
        else:
            red_next=None
        a[:,k]=v
        for c in range(k+1,B): a[:,c]-=s[c]*v; rk[c]-=s[c]
        Rn[k,k]=beta; Rn[k,k+1:]=rk[k+1:]
        # note: R rows k'>k untouched
    return Rn,nrecomp
def test(a,Rjj):
    G=Rjj.T@Rjj+a.T@a
    out=[]
    for pair in (False,True):
        Rn,nr=panel(Rjj,a,pair)
        out.append((np.abs(Rn.T@Rn-G).max()/np.abs(G).max(),nr))
    return out
Rjj=np.triu(rng.standard_normal((8,8)))
print("random", test(rng.standard_normal((64,8)),Rjj))
a=rng.standard_normal((64,8)); a[:,3]=a[:,2]+1e-9*rng.standard_normal(64); Rz=np.zeros((8,8))
print("near-dependent cols, R=0", test(a,Rz))
a=rng.standard_normal((64,3))@rng.standard_normal((3,8))
print("rank-3 block, R=0", test(a,Rz))
a=rng.standard_normal((64,8))*np.logspace(0,-14,8)[None,:]
print("graded", test(a,Rjj*np.logspace(0,-14,8)[None,:]))
a=1e-12*rng.standard_normal((64,8))
print("tiny A vs R", test(a,Rjj))
