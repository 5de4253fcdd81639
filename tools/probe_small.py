"""Per-iteration cost of small (launch-bound) problems with and without the CUDA-graph replay: python tools/probe_small.py"""
import sys, time, json, os
import numpy as np
sys.path.insert(0, ".")
import vbmf_b200_loader
vb = vbmf_b200_loader.load()
ctx = vb.Context(0)
for (L, M, H, kind) in [(10, 20, 2, "dense"), (38, 160, 20, "dual"), (200, 2000, 16, "sparse")]:
    rng = np.random.default_rng(0)
    Y = np.asfortranarray(rng.standard_normal((L, 3)) @ rng.standard_normal((3, M)) + 0.1 * rng.standard_normal((L, M)))
    ctx.attach(Y, force=True)
    if kind == "dense":
        p = vb.vbmf_init(Y, H, rng=rng); fl = vb._lib.EST_COVS | vb._lib.EST_VAR
    elif kind == "dual":
        p = vb.vbmf_dual_init(Y, H, H - 1, rng=rng); fl = vb._lib.EST_CB | vb._lib.EST_PRIORS | vb._lib.FULL_COV
    else:
        p = vb.vbmf_sparse_init(Y, H, rng=rng); fl = vb._lib.EST_CB
    s = vb.Solver(ctx, p); s.upload(p)
    s.run(20, eps=0.0, flags=fl); ctx.sync()
    n = 400
    t = time.perf_counter(); it, d = s.run(n, eps=0.0, flags=fl); ctx.sync(); dt = time.perf_counter() - t
    print(json.dumps({"L": L, "M": M, "H": H, "kind": kind, "graph": os.environ.get("VBMF_B200_NO_GRAPH") is None, "iters": it, "us_per_iter": dt / n * 1e6}))
    s.close()
