"""CPU, world_size-2 (gloo) coverage of the N>1 path's host logic and of the sharding decomposition itself:

every rank owns a column shard of Y and the matching rows of AHat / per-element vectors, computes the local partials
[Y_g*A_g | A_g'A_g | sum of Sigma blocks | dual sums], one all-reduce of that packed buffer per iteration (the same payload
layout libvbmf_b200 sends through NCCL, csrc/kernels.cuh packed_*), then every rank repeats the replicated B-side update.
The result must equal the unsharded oracle -- including Quirk Q2, whose source index depends on the GLOBAL column index.
This is test infrastructure (oracle arithmetic + gloo); the product path has no CPU mode.
"""
import copy
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import vbmf_oracle as vo  # noqa: E402
from tests.helpers import synth  # noqa: E402


def shard_columns(M, world, rank):
    import vbmf_b200_loader
    return vbmf_b200_loader.load().shard_columns(M, world, rank)


def _sharded_sparse_iteration(Y, p, off, n, M, full_cov, est_cb, dual, est_priors, exchange="allreduce", world=1, rank=0):
    """One iteration on the shard [off, off+n): local A side, the exchange, replicated B side.  exchange = "allreduce": one
    packed all-reduce (the NCCL path); "px": the peer-exchange data flow -- small sums gathered and added in rank order,
    Y*AHat reduce-scattered by the planner's row tiles into a row-sharded BHat update, BHat rows all-gathered, tr(BHat.*Q) and
    BHat'BHat as rank partials."""
    H, L = p.H, p.L
    Yg = Y[:, off:off + n]
    # ---- updateA! on the shard (src/vbmf_sparse.jl:176-247): Q2 uses the global index j = (off+m)*H + h
    P = Yg.T @ p.BHat
    CA = p.CA.reshape(n, H)
    if full_cov:
        G = p.sigmaHat * (p.BHat.T @ p.BHat + L * p.SigmaB)
        A = np.empty((n, H)); s = np.empty((n, H)); SA = np.zeros((H, H))
        for m in range(n):
            Sm = np.linalg.inv(G + np.diag(CA[m]))
            A[m] = (p.sigmaHat * Sm) @ P[m]; s[m] = np.diag(Sm); SA += Sm
    else:
        d = vo._diag_precision_base(p, False)
        j0 = (off * H + np.arange(n * H)).astype(np.int64)
        src = np.where(j0 < H, j0, (j0 - H) // max(M - 1, 1))
        s = (1.0 / (d[src] + p.CA)).reshape(n, H)
        A = (p.sigmaHat * s) * P
        SA = np.diag(s.sum(axis=0))
    p.AHat, p.ATVecHat, p.diagSigmaATVec = A, A.reshape(-1).copy(), s.reshape(-1).copy()
    # ---- updateCA! (local; hoisted before the exchange, it does not touch BHat)
    extras = np.zeros(8)
    if dual:
        p.alpha0, p.alpha1 = p.alpha00 + 0.5, p.alpha01 + 0.5
        g1 = (np.arange(H) >= p.H0)[None, :].repeat(n, 0)
        beta = np.where(g1, p.beta01, p.beta00) + 0.5 * (A * A + s)
        ca = np.where(g1, p.alpha1, p.alpha0) / beta
        extras[:4] = [ca[~g1].sum(), ca[g1].sum(), np.log(beta[~g1]).sum(), np.log(beta[g1]).sum()]
        p.beta, p.CA = beta.reshape(-1), ca.reshape(-1)
    else:
        p.beta = p.beta0 + 0.5 * (p.ATVecHat ** 2 + p.diagSigmaATVec)
        p.CA = p.alpha / p.beta
    if exchange == "px":
        small = sum(_gather(torch.from_numpy(np.concatenate([(A.T @ A).reshape(-1), SA.reshape(-1), extras])), world))
        AtA = small[:H * H].reshape(H, H); p.SigmaA = small[H * H:2 * H * H].reshape(H, H).copy(); ex = small[-8:]
        GA = AtA + p.SigmaA
        p.SigmaB = np.linalg.inv(np.diag(p.CB) + p.sigmaHat * GA)
        Qall = _gather(torch.from_numpy(np.ascontiguousarray(Yg @ A)), world)
        lo, hi = _px_rows(L, world, rank)
        Qrows = sum(q[lo:hi] for q in Qall)
        Brows = (p.sigmaHat * Qrows) @ p.SigmaB
        pad = torch.zeros((-(-L // world) + 32, H), dtype=torch.float64)
        pad[:hi - lo] = torch.from_numpy(Brows)
        Bnew = np.empty((L, H))
        for r, rows in enumerate(_gather(pad, world)):
            rlo, rhi = _px_rows(L, world, r)
            Bnew[rlo:rhi] = rows[:rhi - rlo]
        tot = sum(_gather(torch.from_numpy(np.concatenate([(Brows.T @ Brows).reshape(-1), [float(np.sum(Brows * Qrows))]])), world))
        p.BHat, BtB, trBQ = Bnew, tot[:H * H].reshape(H, H), float(tot[-1])
    else:
        # ---- packed all-reduce: [Q | A'A | SA | extras]
        packed = np.concatenate([(Yg @ A).reshape(-1), (A.T @ A).reshape(-1), SA.reshape(-1), extras])
        t = torch.from_numpy(packed)
        dist.all_reduce(t)
        Q = packed[:L * H].reshape(L, H); AtA = packed[L * H:L * H + H * H].reshape(H, H)
        p.SigmaA = packed[L * H + H * H:L * H + 2 * H * H].reshape(H, H).copy(); ex = packed[-8:]
        # ---- replicated: updateB!
        GA = AtA + p.SigmaA
        p.SigmaB = np.linalg.inv(np.diag(p.CB) + p.sigmaHat * GA)
        p.BHat = (p.sigmaHat * Q) @ p.SigmaB
        BtB, trBQ = p.BHat.T @ p.BHat, float(np.sum(p.BHat * Q))
    # ---- replicated: updateCB!, updateSigma!, priors
    if est_cb:
        vo.sparse_updateCB(p)
    p.zeta = p.zeta0 + 0.5 * p.trYTY - trBQ + 0.5 * float(np.sum(GA * (BtB + L * p.SigmaB)))
    p.sigmaHat = p.eta / p.zeta
    if dual and est_priors:
        from scipy.special import digamma
        N0, N1 = M * p.H0, M * p.H1
        for g, (N, a_g, slb) in enumerate(((N0, p.alpha0, ex[2]), (N1, p.alpha1, ex[3]))):
            b0x = p.beta00 if g == 0 else p.beta01
            S = N * digamma(a_g) - slb
            c = N * np.log(b0x)
            try:
                root = vo.fzero_bisect(lambda x: c - N * digamma(x) + S)
                if g == 0:
                    p.alpha00 = root
                else:
                    p.alpha01 = root
            except Exception:
                pass
        p.beta00 = N0 * p.alpha00 / ex[0]
        p.beta01 = N1 * p.alpha01 / ex[1]


def _worker(rank, world, port, case, out, exchange="allreduce"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L, M, H = (24, 101, 4) if exchange == "allreduce" else (150, 101, 4)     # px: five 32-row tiles, uneven shares
    Y = synth(L, M, 2, seed=3)
    off, n = shard_columns(M, world, rank)
    kind, full_cov = case
    dual = kind == "dual"
    pg = vo.vbmf_dual_init(Y, H, 1, rng=np.random.default_rng(7)) if dual else vo.vbmf_sparse_init(Y, H, rng=np.random.default_rng(7))
    p = copy.deepcopy(pg)
    p.M, p.MH = n, n * H
    sl = slice(off * H, (off + n) * H)
    for f in ("ATVecHat", "diagSigmaATVec", "CA", "beta"):
        setattr(p, f, getattr(pg, f)[sl].copy())
    p.AHat = pg.AHat[off:off + n].copy()
    niter = 5
    for _ in range(niter):
        _sharded_sparse_iteration(Y, p, off, n, M, full_cov, True, dual, True, exchange, world, rank)
    if dual:
        vo.vbmf_dual_run(Y, pg, niter, eps=0.0, full_cov=full_cov, est_priors=True, est_cb=True)
    else:
        vo.vbmf_sparse_run(Y, pg, niter, eps=0.0, full_cov=full_cov, est_cb=True)

    def rel(a, b):
        return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
    errs = [rel(p.BHat, pg.BHat), rel(p.SigmaB, pg.SigmaB), rel(p.SigmaA, pg.SigmaA), rel(p.AHat, pg.AHat[off:off + n]),
            rel(p.CA, pg.CA[sl]), abs(p.sigmaHat - pg.sigmaHat) / pg.sigmaHat, rel(p.CB, pg.CB)]
    if dual:
        errs += [abs(p.alpha00 - pg.alpha00) / pg.alpha00, abs(p.beta01 - pg.beta01) / pg.beta01]
    t = torch.tensor([max(errs)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        with open(out, "w") as f:
            f.write(repr(float(t.item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("case", [("sparse", False), ("sparse", True), ("dual", False), ("dual", True)])
def test_world2_sharding_matches_unsharded(case, tmp_path):
    out = str(tmp_path / "err.txt")
    port = 29600 + abs(hash(case)) % 300
    mp.spawn(_worker, args=(2, port, case, out), nprocs=2, join=True)
    assert float(open(out).read()) < 1e-11


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", [("sparse", False), ("dual", True)])
def test_peer_exchange_data_flow_sparse_kinds(case, world, tmp_path):
    """The same decomposition for the sparse / dual kinds: the small part also carries the sum of the Sigma blocks and the
    group sums of the ARD update (hyper-prior roots of vbmf_dual included)."""
    out = str(tmp_path / "err.txt")
    port = 29300 + abs(hash(case)) % 300 + world
    mp.spawn(_worker, args=(world, port, case, out, "px"), nprocs=world, join=True)
    assert float(open(out).read()) < 1e-11


def test_shard_columns_partition():
    for M in (1, 7, 100, 200000):
        for w in (1, 2, 3, 8):
            parts = [shard_columns(M, w, r) for r in range(w)]
            assert parts[0][0] == 0 and sum(n for _, n in parts) == M
            for (o1, n1), (o2, _) in zip(parts, parts[1:]):
                assert o1 + n1 == o2


# ------------------------------------------------------------------------------------------------ peer-exchange data flow
def _px_rows(L, world, rank):
    """This rank's rows of BHat in the peer exchange: the 32-row tiles the library's own host planner assigns it."""
    import ctypes as C
    import vbmf_b200_loader
    lib = vbmf_b200_loader.load()._lib.load()
    out = (C.c_int64 * 3)()
    assert lib.vbmf_b200_px_plan(L, world, rank, C.cast(out, C.c_void_p)) == 0
    return min(out[0] * 32, L), min(out[1] * 32, L)


def _gather(x, world):
    parts = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(parts, x)
    return [t.numpy() for t in parts]


def _px_dense_iteration(Y, p, off, n, M, world, rank):
    """One dense iteration with the exchange the device runs for world > 1 (csrc/kernels.cu, peer exchange): local A side; the
    small sums gathered from every rank and added in rank order; every rank reduces ITS rows of Y*AHat over the ranks, runs
    the BHat epilogue on them and broadcasts the rows; the Grams of BHat / BHat - Bold travel as rank partials."""
    L, H = p.L, p.H
    Yg = Y[:, off:off + n]
    norm_old = float(np.sqrt(np.linalg.eigvalsh(p.BtB)[-1]))          # norm(old), carried between iterations on the device
    # updateA! on the shard (src/vbmf.jl:95-102); SigmaA is replicated
    p.SigmaA = p.sigma2 * vo._inv(p.BtB + L * p.SigmaB + p.sigma2 * p.invCA)
    A = ((Yg.T @ p.BHat) @ p.SigmaA) / p.sigma2
    p.AHat = A
    Qloc = Yg @ A
    # small exchange: every rank's A'A, summed in rank order
    AtA = sum(_gather(torch.from_numpy(np.ascontiguousarray(A.T @ A)), world))
    p.SigmaB = p.sigma2 * vo._inv(AtA + M * p.SigmaA + p.sigma2 * p.invCB)
    # reduce-scatter of Y*AHat into the row-sharded epilogue
    Qall = _gather(torch.from_numpy(np.ascontiguousarray(Qloc)), world)
    lo, hi = _px_rows(L, world, rank)
    Qrows = sum(q[lo:hi] for q in Qall)
    Brows = (Qrows @ p.SigmaB) / p.sigma2
    Drows = Brows - p.BHat[lo:hi]
    part = np.concatenate([(Brows.T @ Brows).reshape(-1), (Drows.T @ Drows).reshape(-1), [float(np.sum(Brows * Qrows))]])
    # all-gather of the BHat rows (every rank writes its rows into every peer's BHat) and of the rank partials
    Bnew = np.empty_like(p.BHat)
    pad = torch.zeros((-(-L // world) + 32, H), dtype=torch.float64)
    pad[:hi - lo] = torch.from_numpy(Brows)
    for r, rows in enumerate(_gather(pad, world)):
        rlo, rhi = _px_rows(L, world, r)
        Bnew[rlo:rhi] = rows[:rhi - rlo]
    parts = _gather(torch.from_numpy(part), world)
    tot = sum(parts)
    p.BHat = Bnew
    p.BtB, DtD, trBQ = tot[:H * H].reshape(H, H), tot[H * H:2 * H * H].reshape(H, H), tot[-1]
    # replicated tail from the exchanged Grams: updateCA!, updateCB!, updateSigma2!, delta (src/vbmf.jl:129-157, src/util.jl:27-29)
    for h in range(H):
        p.CA[h, h] = AtA[h, h] / M + p.SigmaA[h, h]
        p.CB[h, h] = p.BtB[h, h] / L + p.SigmaB[h, h]
    p.invCA, p.invCB = vo._inv(p.CA), vo._inv(p.CB)
    p.sigma2 = (p.trYTY - 2.0 * trBQ + float(np.trace((AtA + M * p.SigmaA) @ (p.BtB + L * p.SigmaB)))) / (L * M)
    return float(np.sqrt(np.linalg.eigvalsh(DtD)[-1])) / norm_old      # delta = norm(old - new) / norm(old), spectral (Q1)


def _px_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L, M, H = 150, 211, 6            # 5 row tiles of 32 (the last one ragged): uneven shares between the ranks
    Y = synth(L, M, 3, seed=4)
    off, n = shard_columns(M, world, rank)
    pg = vo.vbmf_init(Y, H, ca=1.0, cb=1.0, sigma2=1.0, rng=np.random.default_rng(8))
    p = copy.deepcopy(pg)
    p.AHat = pg.AHat[off:off + n].copy()
    p.BtB = p.BHat.T @ p.BHat
    p.trYTY = float(np.sum(Y * Y))
    niter = 6
    d = None
    for _ in range(niter):
        d = _px_dense_iteration(Y, p, off, n, M, world, rank)
    _, it_o, d_o = vo.vbmf_run(Y, pg, niter, eps=0.0, est_covs=True, est_var=True, yhat=False)

    def rel(a, b):
        return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
    errs = [rel(p.BHat, pg.BHat), rel(p.SigmaB, pg.SigmaB), rel(p.SigmaA, pg.SigmaA), rel(p.AHat, pg.AHat[off:off + n]),
            rel(p.CA, pg.CA), rel(p.CB, pg.CB), abs(p.sigma2 - pg.sigma2) / pg.sigma2, abs(d - d_o) / d_o * 1e-2]
    t = torch.tensor([max(errs)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        with open(out, "w") as f:
            f.write(repr(float(t.item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_exchange_data_flow_matches_unsharded(world, tmp_path):
    """The decomposition the device uses instead of an all-reduce (small sums in rank order, reduce-scatter of Y*AHat by the
    planner's row tiles, row-sharded BHat epilogue, all-gather of the rows, Grams as rank partials) reproduces the unsharded
    oracle, with uneven row shares (world 3: 2 / 1 / 2 tiles, the last tile ragged)."""
    out = str(tmp_path / "err.txt")
    mp.spawn(_px_worker, args=(world, 29950 + world, out), nprocs=world, join=True)
    assert float(open(out).read()) < 1e-11
