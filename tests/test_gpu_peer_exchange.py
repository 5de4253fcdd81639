"""Peer exchange (updateB!'s reduce-scatter / row-sharded epilogue / all-gather over peer-mapped memory, kernels.cuh PxDev) on ONE
GPU: VBMF_B200_PX_SELF=1 makes a single rank run the exchange kernels against itself (same kernels, W = 1), so tile ranges,
the in-place row sums, the rank-partial slots and the barrier epochs are exercised without a second device.  With one rank
every sum has one term, so the results must equal the ordinary path BIT FOR BIT, and match the oracle at 1e-10.
The W > 1 runs are tests/test_gpu_multi.py (2 / 4 / 8 GPUs) and tests/mgpu_worker.py."""
import copy

import numpy as np
import pytest

from oracle import vbmf_oracle as vo
from tests.helpers import synth

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def G():
    from tests import gpu_helpers
    return gpu_helpers


def _ctx(G, monkeypatch, px):
    if px:
        monkeypatch.setenv("VBMF_B200_PX_SELF", "1")
    else:
        monkeypatch.delenv("VBMF_B200_PX_SELF", raising=False)
    return G.vb.Context(device=0)


def _run(G, ctx, kind, Y, p, niter, kw):
    q = G.to_gpu_params(p)
    Yf = np.asfortranarray(Y)
    if kind == "dense":
        G.vb.vbmf_(Yf, q, niter, eps=0.0, ctx=ctx, **kw)
    elif kind == "sparse":
        G.vb.vbmf_sparse_(Yf, q, niter, eps=0.0, ctx=ctx, **kw)
    else:
        G.vb.vbmf_dual_(Yf, q, niter, eps=0.0, ctx=ctx, **kw)
    return q


CASES = [
    # kind, L, M, H, kwargs
    ("dense", 96, 1001, 8, dict(est_covs=True, est_var=True)),
    ("dense", 1000, 3001, 64, dict(est_covs=True, est_var=True)),          # 32 row tiles, last one ragged
    ("dense", 33, 200, 5, dict(est_covs=True, est_var=False)),            # odd H (SIMT K2), two tiles
    ("sparse", 200, 900, 32, dict(full_cov=False, est_cb=True)),
    ("sparse", 70, 300, 12, dict(full_cov=True, est_cb=True)),
    ("dual", 130, 640, 16, dict(full_cov=False, est_priors=True, est_cb=True)),
    ("sparse", 64, 333, 8, dict(full_cov=False, diag_var=True, est_cb=True)),   # heteroscedastic: stays on the all-reduce path
]


@pytest.mark.parametrize("kind,L,M,H,kw", CASES)
def test_self_exchange_bit_identical_and_parity(G, monkeypatch, kind, L, M, H, kw):
    Y = synth(L, M, max(1, H // 2), seed=L + M)
    rng = np.random.default_rng(3)
    if kind == "dense":
        p = vo.vbmf_init(Y, H, ca=1.0, cb=1.0, sigma2=1.0, rng=rng)
    elif kind == "sparse":
        p = vo.vbmf_sparse_init(Y, H, rng=rng)
    else:
        p = vo.vbmf_dual_init(Y, H, H // 2, rng=rng)
    res = {}
    for px in (False, True):
        ctx = _ctx(G, monkeypatch, px)
        try:
            res[px] = [_run(G, ctx, kind, Y, p, n, kw) for n in (1, 5)]
            if px:
                assert ctx.peer_exchange()
            else:
                assert not ctx.peer_exchange()
        finally:
            ctx.close()
    for a, b in zip(res[False], res[True]):
        for f in G.FIELDS[kind]:
            assert np.array_equal(np.asarray(getattr(a, f)), np.asarray(getattr(b, f))), f
        assert a.iterations == b.iterations and getattr(a, "d", None) == getattr(b, "d", None)
    # one iteration from the oracle's state (north_star's per-iteration bar)
    po = copy.deepcopy(p)
    if kind == "dense":
        vo.vbmf_run(Y, po, 1, eps=0.0, **kw)
    elif kind == "sparse":
        vo.vbmf_sparse_run(Y, po, 1, eps=0.0, **kw)
    else:
        vo.vbmf_dual_run(Y, po, 1, eps=0.0, **kw)
    G.compare(res[True][0], po, TOL)


def test_self_exchange_repeated_calls_and_early_exit(G, monkeypatch):
    """Barrier epochs keep growing across runs and solvers of one context; iterations enqueued past convergence are skipped
    as a whole (no barrier is entered), and the iteration count matches the oracle's."""
    ctx = _ctx(G, monkeypatch, True)
    try:
        Y = synth(50, 300, 2, seed=11)
        p = vo.vbmf_init(Y, 4, ca=1.0, cb=1.0, sigma2=1.0, rng=np.random.default_rng(12))
        Yf = np.asfortranarray(Y)
        for eps in (1e-3, 1e-4):
            po = copy.deepcopy(p)
            _, it_o, d_o = vo.vbmf_run(Y, po, 500, eps=eps, est_covs=True, est_var=True)
            q = G.to_gpu_params(p)
            G.vb.vbmf_(Yf, q, 500, eps=eps, est_covs=True, est_var=True, ctx=ctx)
            assert it_o < 500 and q.iterations == it_o
            assert abs(q.d - d_o) <= 1e-7 * d_o
        assert ctx.peer_exchange()
    finally:
        ctx.close()
