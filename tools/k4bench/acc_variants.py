import numpy as np, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools/k4bench')
from oracle import vbmf_oracle as vo
from tests.helpers import synth
from acc_compare import gj_ld, gj_rank1
L,M,H=100,257,16
Y=synth(L,M,8,seed=L+M)
p=vo.vbmf_init(Y,H,rng=np.random.default_rng(L+M+1))
for it in range(10): vo.vbmf_run(Y,p,1,eps=0.0,est_covs=False,est_var=True)
S = p.BHat.T@p.BHat + p.L*p.SigmaB + p.sigma2*p.invCA
P=Y.T@p.BHat
sc=1/np.sqrt(np.diag(S)); Se=S*sc[:,None]*sc[None,:]
true=(gj_ld(Se).astype(np.float64))
rel=lambda a,b: np.max(np.abs(a-b))/np.max(np.abs(b))
def block_gj(Se, bs, trick):
    A=Se.copy(); n=A.shape[0]
    for s in range(0,n,bs):
        K=np.arange(s,min(n,s+bs)); R=np.setdiff1d(np.arange(n),K)
        Pn=A[:,K].copy(); D=A[np.ix_(K,K)].copy()
        # sequential sweeps in the panel
        X=Pn.copy()
        for c,kc in enumerate(K):
            pr=X[kc].copy(); idv=1.0/pr[c]
            for l in range(n):
                f=(1.0-idv) if l==kc else X[l,c]*idv
                for q in range(len(K)):
                    if q!=c: X[l,q]-=f*pr[q]
                X[l,c]=-idv if l==kc else f
        if trick:
            Wp=X.copy(); Pp=Pn.copy()
            for c,kc in enumerate(K): Wp[kc,c]+=1.0; Pp[kc,c]-=1.0
            A=A-Wp@Pp.T
            for kc in K: A[kc,kc]-=2.0
        else:
            A[np.ix_(R,R)]-=X[R]@Pn[R].T
            A[:,K]=X; A[K,:]=X.T
    return -A
for name,inv in (("numpy",np.linalg.inv(Se)),("rank1",gj_rank1(Se)),("block4 trick",block_gj(Se,4,True)),("block4 overwrite",block_gj(Se,4,False)),("block1 trick",block_gj(Se,1,True)),("block1 overwrite",block_gj(Se,1,False))):
    X=inv*sc[:,None]*sc[None,:]; T=true*sc[:,None]*sc[None,:]
    print("%-18s inv err %.1e   AHat err %.1e   resid |S X - I| %.1e"%(name, rel(X,T), rel(P@X,P@T), np.max(np.abs(S@X-np.eye(H)))))
def block_gj_seqterms(Se, bs):
    """rank-bs update built from the SEQUENTIAL sweeps' own terms: W'[:,c] = multipliers of sweep c, P'[:,c] = column c right
    before its sweep (minus e on the pivot entry): S <- S - W'P'^T - 2E (the products a rank-1 sweep would apply one by one)."""
    A=Se.copy(); n=A.shape[0]
    for s in range(0,n,bs):
        K=np.arange(s,min(n,s+bs))
        X=A[:,K].copy(); Wp=np.zeros_like(X); Pp=np.zeros_like(X)
        for c,kc in enumerate(K):
            pr=X[kc].copy(); idv=1.0/pr[c]
            Pp[:,c]=X[:,c]; Pp[kc,c]-=1.0
            for l in range(n):
                f=(1.0-idv) if l==kc else X[l,c]*idv
                Wp[l,c]=f
                for q in range(len(K)):
                    if q!=c: X[l,q]-=f*pr[q]
                X[l,c]=-idv if l==kc else f
        A=A-Wp@Pp.T
        for kc in K: A[kc,kc]-=2.0
    return -A
for bs in (4,8):
    inv=block_gj_seqterms(Se,bs)
    X=inv*sc[:,None]*sc[None,:]; T=true*sc[:,None]*sc[None,:]
    print("block%d seq-terms   inv err %.1e   AHat err %.1e   resid %.1e"%(bs, rel(X,T), rel(P@X,P@T), np.max(np.abs(S@X-np.eye(H)))))
