import sys
import numpy as np
sys.path.insert(0, ".")
import vbmf_b200_loader
vb = vbmf_b200_loader.load()
from tests import gpu_helpers as G
from oracle import vbmf_oracle as vo
import copy
for (L, M, H) in [(64, 200000, 64), (200, 100000, 32), (3000, 50000, 64)]:
    rng = np.random.default_rng(0)
    Y = np.asfortranarray(rng.standard_normal((L, M)))
    B = np.asfortranarray(rng.standard_normal((L, H)))
    A = np.asfortranarray(rng.standard_normal((M, H)))
    ctx = vb.Context(0)
    ctx.attach(Y)
    P = ctx.gemm_YtB(B); Q = ctx.gemm_YA(A)
    eP = np.abs(P - Y.T @ B).max(axis=1) / np.abs(Y.T @ B).max()
    print(L, M, H, "K1 relerr %.2e  K2 relerr %.2e" % (eP.max(), G.rel(Q, Y @ A)), "bad rows:", np.where(eP > 1e-12)[0][:10], (eP > 1e-12).sum())
    p = vo.vbmf_init(Y, H, rng=np.random.default_rng(1))
    q = G.to_gpu_params(p)
    vo.vbmf_run(Y, p, 1, eps=0.0, est_covs=True, est_var=True)
    vb.vbmf_(Y, q, 1, eps=0.0, est_covs=True, est_var=True, ctx=ctx)
    errs = {f: G.rel(getattr(q, f), getattr(p, f)) for f in G.FIELDS["dense"]}
    print("   loop 1 iter:", {k: "%.1e" % v for k, v in errs.items()})
    eA = np.abs(q.AHat - p.AHat).max(axis=1) / np.abs(p.AHat).max()
    bad = np.where(eA > 1e-10)[0]
    print("   bad A rows:", bad[:20], len(bad), "tiles:", sorted(set((bad // 128).tolist()))[:20])
    ctx.close()
