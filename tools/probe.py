"""Quick timing probe of the resident-state loop (not the bench contract): python tools/probe.py L M H kind niter"""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import vbmf_b200_loader
vb = vbmf_b200_loader.load()

L, M, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kind = sys.argv[4] if len(sys.argv) > 4 else "dense"
niter = int(sys.argv[5]) if len(sys.argv) > 5 else 10
flagstr = sys.argv[6] if len(sys.argv) > 6 else ""
ctx = vb.Context(0)
t0 = time.time(); ctx.synth(L, M, rank=H // 2, noise=0.1, seed=20260101); ctx.sync(); t_synth = time.time() - t0
rng = np.random.default_rng(1)
Yshape = np.empty((0, 0))
class FakeY:  # init only needs the shape
    shape = (L, M)
if kind == "dense":
    p = vb.vbmf_init(FakeY, H, rng=rng)
    flags = vb._lib.EST_COVS | vb._lib.EST_VAR
elif kind == "sparse":
    p = vb.vbmf_sparse_init(FakeY, H, rng=rng, trYTY=ctx.trYTY())
    flags = vb._lib.EST_CB | (vb._lib.FULL_COV if "full" in flagstr else 0) | (vb._lib.DIAG_VAR if "dv" in flagstr else 0)
else:
    p = vb.vbmf_dual_init(FakeY, H, H // 2, rng=rng, trYTY=ctx.trYTY())
    flags = vb._lib.EST_CB | vb._lib.EST_PRIORS | (vb._lib.FULL_COV if "full" in flagstr else 0)
s = vb.Solver(ctx, p)
s.upload(p)
it, d = s.run(3, eps=0.0, flags=flags)
ctx.sync()
ctx.profile(True)
n0 = vb._lib.load().vbmf_b200_launch_count()
t0 = time.time(); it, d = s.run(niter, eps=0.0, flags=flags); ctx.sync(); t = time.time() - t0
prof = ctx.profile_read()
n1 = vb._lib.load().vbmf_b200_launch_count()
flops = 4.0 * L * M * H
out = {"L": L, "M": M, "H": H, "kind": kind, "flags": flagstr, "iters": it, "d": d, "ms_per_iter": t / niter * 1e3, "it_per_s": niter / t,
       "tflops": flops * niter / t * 1e-12, "frac_of_37.0": flops * niter / t / 37.0e12,
       "k1_ms": prof["k1_ms"] / max(prof["k1_launches"], 1), "k2_ms": prof["k2_ms"] / max(prof["k2_launches"], 1),
       "k1_tflops": 2.0 * L * M * H / (prof["k1_ms"] / max(prof["k1_launches"], 1)) * 1e-9,
       "k2_tflops": 2.0 * L * M * H / (prof["k2_ms"] / max(prof["k2_launches"], 1)) * 1e-9,
       "launches_per_iter": (n1 - n0) / niter, "synth_s": t_synth}
print(json.dumps(out))
