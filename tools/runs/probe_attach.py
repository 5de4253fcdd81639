"""Probe: upload time of Y through attach_Y (synchronous vs chunked asynchronous) and the first consumer's wait."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vbmf_b200_loader
vb = vbmf_b200_loader.load()
L, M, H = 10000, 100000, 32
ypin = torch.empty((M, L), dtype=torch.float64, pin_memory=True)
Yh = ypin.numpy().T
Yh[...] = 0.5
ctx = vb.Context(device=0)
for mode in ("0", "2048", "512", "0", "2048"):
    os.environ["VBMF_B200_ATTACH_CHUNK_MB"] = mode
    t0 = time.perf_counter(); ctx.attach(Yh); t1 = time.perf_counter(); ctx.sync(); t2 = time.perf_counter()
    print("chunk_mb=%s attach returns after %.3f s, upload done after %.3f s (%.1f GB/s)" % (mode, t1 - t0, t2 - t0, L * M * 8 / 1e9 / (t2 - t0)), flush=True)
    t0 = time.perf_counter(); ctx.attach(Yh); t1 = time.perf_counter(); tr = ctx.trYTY(); t2 = time.perf_counter()
    print("   attach + trYTY(): %.3f s  tr=%.6g" % (t2 - t0, tr), flush=True)
    rng = np.random.default_rng(0)
    class Shape: shape = (L, M)
    p = vb.vbmf_sparse_init(Shape, H, rng=rng, trYTY=tr)
    t0 = time.perf_counter(); ctx.attach(Yh); t1 = time.perf_counter()
    vb.vbmf_sparse_(None, p, 20, eps=0.0, ctx=ctx, yhat=False); t2 = time.perf_counter()
    print("   attach %.3f s + vbmf_sparse_(20 it) %.3f s = %.3f s" % (t1 - t0, t2 - t1, t2 - t0), flush=True)
