"""Accuracy of the blocked (rank-4, DMMA-shaped) Gauss-Jordan inverse vs the rank-1 sweep it replaces and vs LAPACK's LU
inverse, all against a long-double reference, on equilibrated ARD-like precision matrices.  Test infrastructure only."""
import numpy as np
from emulate_blockgj import warp_block_gj
def gj_ld(S):
    A=S.astype(np.longdouble).copy(); n=A.shape[0]
    for k in range(n):
        d=A[k,k]; col=A[:,k].copy(); row=A[k,:].copy()
        A=A-np.outer(col,row)/d
        A[:,k]=col/d; A[k,:]=row/d; A[k,k]=-1/d
    return (-A)
def gj_rank1(S):
    A=S.copy(); n=A.shape[0]
    for k in range(n):
        c=A[:,k].copy(); d=c[k]; idv=1.0/d
        for l in range(n):
            t=c[l]*idv
            beta=(idv-1.0) if l==k else -t
            col=A[:,l]+c*beta
            col[k]=-idv if l==k else t
            A[:,l]=col
    return -A
if __name__ == "__main__":
    rng=np.random.default_rng(1)
    res=[]
    for trial in range(300):
        N=32
        L=rng.integers(3,40)
        Bm=rng.standard_normal((L,N))*rng.uniform(0.01,10,(1,N))
        ca=10.0**rng.uniform(-3,10,N)
        S=rng.uniform(0.1,100)*(Bm.T@Bm)+np.diag(ca)
        sc=1/np.sqrt(np.diag(S)); Se=S*sc[:,None]*sc[None,:]
        ref=(gj_ld(Se)).astype(np.float64)
        mx=lambda X: np.max(np.abs(X-ref))/np.max(np.abs(ref))
        a=np.linalg.inv(Se); b=gj_rank1(Se); c=-warp_block_gj(Se,4)
        res.append((np.linalg.cond(Se), mx(a), mx(b), mx(c)))
    res=np.array(res)
    print("cond max", res[:,0].max())
    for i,n in enumerate(["numpy","rank1","block4"]):
        print(n, "max rel(max-norm) err", res[:,1+i].max(), "median", np.median(res[:,1+i]), "max err/cond", (res[:,1+i]/res[:,0]).max())
