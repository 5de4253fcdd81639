"""Full-size check: oracle update loop fed with the GPU's contractions vs the resident GPU loop, iteration by iteration."""
import sys, copy
import numpy as np
sys.path.insert(0, ".")
import vbmf_b200_loader
vb = vbmf_b200_loader.load()
from oracle import vbmf_oracle as vo
from tests import gpu_helpers as G
L, M, H, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ctx = vb.Context(0)
ctx.synth(L, M, rank=H // 2, noise=0.1, seed=20260101)
Yc = vo.ContractedY((L, M), ctx.gemm_YtB, ctx.gemm_YA, ctx.trYTY())
po = vo.vbmf_init(Yc, H, rng=np.random.default_rng(1))
q = G.to_gpu_params(po)
s = vb.Solver(ctx, q); s.upload(q)
flags = vb._lib.EST_COVS | vb._lib.EST_VAR
for i in range(n):
    vo.vbmf_run(Yc, po, 1, eps=0.0, est_covs=True, est_var=True)
    it, d = s.run(1, eps=0.0, flags=flags)
    s.download(q)
    errs = {f: G.rel(getattr(q, f), getattr(po, f)) for f in G.FIELDS["dense"]}
    print(i + 1, "sigma2 %.6e/%.6e maxCA %.3e/%.3e  maxerr %.2e (%s)" % (q.sigma2, po.sigma2, np.diag(q.CA).max(), np.diag(po.CA).max(),
          max(errs.values()), max(errs, key=errs.get)), flush=True)
