"""Worker for the multi-GPU parity test (launched by torchrun, one rank per GPU): column-sharded solvers vs the oracle
on the gathered problem.  Prints 'MGPU OK' on rank 0."""
import copy
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vbmf_b200_loader  # noqa: E402
from oracle import vbmf_oracle as vo  # noqa: E402
from tests import gpu_helpers as G  # noqa: E402
from tests.helpers import synth  # noqa: E402

vb = vbmf_b200_loader.load()


def shard(pg, off, n, kind):
    p = copy.deepcopy(pg)
    H = p.H
    p.M = n
    p.AHat = np.asfortranarray(pg.AHat[off:off + n])
    if kind != "dense":
        p.MH = n * H
        sl = slice(off * H, (off + n) * H)
        for f in ("ATVecHat", "diagSigmaATVec", "CA", "beta"):
            setattr(p, f, getattr(pg, f)[sl].copy())
    if kind == "dual":
        p.A0Hat = np.asfortranarray(pg.A0Hat[off:off + n]); p.A1Hat = np.asfortranarray(pg.A1Hat[off:off + n])
        for f, w in (("CA0", p.H0), ("beta0", p.H0), ("CA1", p.H1), ("beta1", p.H1)):
            setattr(p, f, getattr(pg, f)[off * w:(off + n) * w].copy())
    if kind != "dual" and len(pg.labels):
        lab = np.asarray(pg.labels)
        p.labels = lab[(lab > off) & (lab <= off + n)] - off
    return p


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(vb.Context.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx = vb.Context(device=local, rank=rank, world=world, nccl_id=bytes(idt.cpu().numpy().tobytes()))
    L, M, H = 96, 1001, 8
    Y = synth(L, M, 4, seed=5)
    off, n = vb.shard_columns(M, world, rank)
    ctx.attach(np.asfortranarray(Y[:, off:off + n]), M_global=M, col_offset=off)
    assert abs(ctx.trYTY() - float(np.sum(Y * Y))) < 1e-10 * float(np.sum(Y * Y))
    worst = 0.0
    cases = [("dense", {}, dict(est_covs=True, est_var=True)),
             ("sparse", {}, dict(full_cov=False, est_cb=True)),            # Q2 map needs the global column index
             ("sparse", {"H1": 2, "labels": list(range(3, M, 7))}, dict(full_cov=True, diag_var=True, est_cb=True)),
             ("dual", {}, dict(full_cov=False, est_priors=True, est_cb=True)),
             ("dual", {}, dict(full_cov=True, est_priors=True, est_cb=True))]
    for kind, ikw, kw in cases:
        rng = np.random.default_rng(11)
        if kind == "dense":
            po = vo.vbmf_init(Y, H, rng=rng, **ikw)
        elif kind == "sparse":
            po = vo.vbmf_sparse_init(Y, H, rng=rng, **ikw)
        else:
            po = vo.vbmf_dual_init(Y, H, 3, rng=rng)
        q = shard(G.to_gpu_params(po), off, n, kind)
        niter = 6
        if kind == "dense":
            _, it_o, d_o = vo.vbmf_run(Y, po, niter, eps=1e-12, **kw)
            vb.vbmf_(None, q, niter, eps=1e-12, ctx=ctx, **kw); d = q.d
            lb = lb_o = 0.0
        elif kind == "sparse":
            d_o, it_o = vo.vbmf_sparse_run(Y, po, niter, eps=1e-12, **kw)
            d = vb.vbmf_sparse_(None, q, niter, eps=1e-12, ctx=ctx, **kw)
            lb_o = vo.sparse_lowerBound(Y, po); lb = vb.lowerBound(None, q, ctx=ctx)
        else:
            d_o, it_o = vo.vbmf_dual_run(Y, po, niter, eps=1e-12, **kw)
            d = vb.vbmf_dual_(None, q, niter, eps=1e-12, ctx=ctx, **kw)
            lb_o = vo.dual_lowerBound(Y, po); lb = vb.lowerBound(None, q, ctx=ctx)
        assert q.iterations == it_o, (kind, q.iterations, it_o)
        ref = shard(po, off, n, kind)
        fields = [f for f in G.FIELDS[kind]]
        errs = {f: G.rel(getattr(q, f), getattr(ref, f)) for f in fields}
        errs["YHat"] = G.rel(q.YHat, (po.BHat @ po.AHat.T)[:, off:off + n])
        errs["d"] = abs(d - d_o) / abs(d_o) * 1e-2
        errs["lb"] = abs(lb - lb_o) / max(abs(lb_o), 1e-300)
        bad = {f: e for f, e in errs.items() if not e < 1e-10}
        assert not bad, (kind, kw, bad)
        worst = max(worst, max(errs.values()))
    # the updateB! exchange ran through peer-mapped memory (own kernels over NVLink) unless it was switched off
    px = ctx.peer_exchange()
    assert px == (world <= 8 and os.environ.get("VBMF_B200_NO_PX") is None), px
    # many row tiles per rank, ragged last tile, H = 64 / 32: dense and sparse (diagonal) through the exchange kernels
    for kind, (L2, M2, H2), kw in (("dense", (1000, 3001, 64), dict(est_covs=True, est_var=True)),
                                   ("sparse", (517, 2000, 32), dict(full_cov=False, est_cb=True)),
                                   ("dual", (300, 1500, 16), dict(full_cov=True, est_priors=True, est_cb=True)),
                                   # L*H large enough that the peer-visible buffer has to grow (collective re-mapping)
                                   ("dense", (40000, 136, 64), dict(est_covs=True, est_var=True))):
        Y2 = synth(L2, M2, 6, seed=17)
        off2, n2 = vb.shard_columns(M2, world, rank)
        ctx.attach(np.asfortranarray(Y2[:, off2:off2 + n2]), M_global=M2, col_offset=off2)
        rng = np.random.default_rng(21)
        po = vo.vbmf_init(Y2, H2, ca=1.0, cb=1.0, sigma2=1.0, rng=rng) if kind == "dense" else (
            vo.vbmf_sparse_init(Y2, H2, rng=rng) if kind == "sparse" else vo.vbmf_dual_init(Y2, H2, 5, rng=rng))
        for it in range(3):                      # teacher-forced: every iteration starts from the oracle's state
            q = shard(G.to_gpu_params(po), off2, n2, kind)
            if kind == "dense":
                vo.vbmf_run(Y2, po, 1, eps=0.0, **kw); vb.vbmf_(None, q, 1, eps=0.0, ctx=ctx, yhat=False, **kw)
            elif kind == "sparse":
                vo.vbmf_sparse_run(Y2, po, 1, eps=0.0, **kw); vb.vbmf_sparse_(None, q, 1, eps=0.0, ctx=ctx, yhat=False, **kw)
            else:
                vo.vbmf_dual_run(Y2, po, 1, eps=0.0, **kw); vb.vbmf_dual_(None, q, 1, eps=0.0, ctx=ctx, yhat=False, **kw)
            ref = shard(po, off2, n2, kind)
            errs = {f: G.rel(getattr(q, f), getattr(ref, f)) for f in G.FIELDS[kind]}
            bad = {f: e for f, e in errs.items() if not e < 1e-10}
            assert not bad, (kind, it, bad)
            worst = max(worst, max(errs.values()))
    t = torch.tensor([worst], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("MGPU OK world=%d worst_rel_err=%.2e peer_exchange=%s" % (world, t.item(), px), flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
