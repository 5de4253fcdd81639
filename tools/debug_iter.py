import sys, json
import numpy as np
sys.path.insert(0, ".")
import vbmf_b200_loader
vb = vbmf_b200_loader.load()
L, M, H, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ctx = vb.Context(0)
ctx.synth(L, M, rank=H // 2, noise=0.1, seed=20260101)
class FakeY: shape = (L, M)
p = vb.vbmf_init(FakeY, H, rng=np.random.default_rng(1))
s = vb.Solver(ctx, p); s.upload(p)
flags = vb._lib.EST_COVS | vb._lib.EST_VAR
for i in range(n):
    it, d = s.run(1, eps=0.0, flags=flags)
    s.download(p)
    print(i + 1, "it", it, "d %.4e sigma2 %.4e minCA %.3e maxCA %.3e minCB %.3e |B| %.3e nanA %d nanB %d SigmaA[0,0] %.3e SigmaB[0,0] %.3e" % (
        d, p.sigma2, np.diag(p.CA).min(), np.diag(p.CA).max(), np.diag(p.CB).min(), np.linalg.norm(p.BHat),
        np.isnan(p.AHat).sum(), np.isnan(p.BHat).sum(), p.SigmaA[0, 0], p.SigmaB[0, 0]), flush=True)
