"""Timing probe of the batched vbls path on a synthetic MIL-like workload (config 2)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import vbmf_b200_loader
vb = vbmf_b200_loader.load()
nbags, L, H, niter = 548, 38, 20, 20
rng = np.random.default_rng(0)
Ys, ps = [], []
B = [np.asfortranarray(rng.standard_normal((L, H))) for _ in range(2)]
for b in range(nbags):
    M = int(rng.integers(5, 41))
    Y = np.asfortranarray(10.0 * (rng.standard_normal((L, 3)) @ rng.standard_normal((3, M)) + 0.1 * rng.standard_normal((L, M))))
    for c in range(2):
        p = vb.vbmf_dual_init(Y, H, H - 1, rng=rng)
        p.BHat = B[c].copy(); p.SigmaB = np.asfortranarray(np.diag(rng.uniform(1e-3, 1e-2, H))); p.sigmaHat = 0.7
        Ys.append(Y); ps.append(p)
ctx = vb.Context(0)
vb.vbls_batched_(Ys[:8], [vb.copy(p) for p in ps[:8]], niter, full_cov=True, ctx=ctx)
for fc in (True, False):
    qs = [vb.copy(p) for p in ps]
    t = time.perf_counter(); vb.vbls_batched_(Ys, qs, niter, full_cov=fc, ctx=ctx); dt = time.perf_counter() - t
    print(json.dumps({"problems": len(qs), "full_cov": fc, "seconds_e2e": dt, "problems_per_s_e2e": len(qs) / dt}))
