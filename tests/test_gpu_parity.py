"""GPU parity tests proper (run with -m gpu on a B200): every path goes through the C ABI (ctypes -> libvbmf_b200.so).

Tolerance: north_star asks for <= 1e-10 relative error on factors / noise / ELBO per iteration and an identical
convergence iteration count.  TOL below is that bar; golden-trajectory checks use it against the reference's own logs.
"""
import copy

import numpy as np
import pytest

from oracle import vbmf_oracle as vo
from tests.helpers import dense_state_from_golden, jl, load_golden, sparse_state_from_golden, synth

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def G():
    from tests import gpu_helpers
    return gpu_helpers


@pytest.fixture(scope="module")
def ctx(G):
    c = G.vb.Context(device=0)
    yield c
    c.close()


# ------------------------------------------------------------------------------------------------ K1 / K2
@pytest.mark.parametrize("L,M,H", [(10, 20, 2), (257, 1000, 64), (1000, 3001, 32), (130, 517, 128), (64, 300, 5),
                                   (2048, 4096, 64), (33, 7, 16), (16, 16, 1),
                                   # persistent CTAs walking many tiles / split-K slabs (cross-proxy WAR regression)
                                   (64, 200000, 64), (3000, 50000, 64), (1000, 100000, 32), (40000, 300, 64), (500, 30000, 128)])
def test_contractions(G, ctx, L, M, H):
    rng = np.random.default_rng(L * 7 + M)
    Y = np.asfortranarray(rng.standard_normal((L, M)))
    B = np.asfortranarray(rng.standard_normal((L, H)))
    A = np.asfortranarray(rng.standard_normal((M, H)))
    ctx.attach(Y, force=True)
    assert abs(ctx.trYTY() - float(np.sum(Y * Y))) <= 1e-12 * float(np.sum(Y * Y))
    P = ctx.gemm_YtB(B)
    Q = ctx.gemm_YA(A)
    # FP64 accumulate in a different order than BLAS: |err| <= ~K*eps*|a||b|
    assert G.rel(P, Y.T @ B) < 1e-13
    assert G.rel(Q, Y @ A) < 1e-13


def test_contraction_linearity_large(G, ctx):
    """Size-independent property at a shape the oracle would not finish quickly: linearity in the small operand."""
    L, M, H = 4096, 20000, 64
    ctx.synth(L, M, rank=8, noise=0.1, seed=7)
    rng = np.random.default_rng(3)
    B1, B2 = rng.standard_normal((L, H)), rng.standard_normal((L, H))
    P1, P2, P12 = ctx.gemm_YtB(B1), ctx.gemm_YtB(B2), ctx.gemm_YtB(B1 + 2.0 * B2)
    assert G.rel(P12, P1 + 2.0 * P2) < 1e-12
    A1, A2 = rng.standard_normal((M, H)), rng.standard_normal((M, H))
    Q1, Q2, Q12 = ctx.gemm_YA(A1), ctx.gemm_YA(A2), ctx.gemm_YA(A1 - 0.5 * A2)
    assert G.rel(Q12, Q1 - 0.5 * Q2) < 1e-12
    # checksum of checksums: sum(P .* A) == sum(Q .* B) == tr(A' Y' B)
    assert abs(np.sum(P1 * A1) - np.sum(Q1 * B1)) <= 1e-11 * abs(np.sum(P1 * A1))


# ------------------------------------------------------------------------------------------------ golden trajectories
@pytest.mark.parametrize("niter", [1, 2, 10, 100])
def test_dense_golden(G, ctx, niter):
    g = load_golden("vbmf_test")
    Y, p = dense_state_from_golden(g)
    q = G.to_gpu_params(p)
    G.vb.vbmf_(np.asfortranarray(Y), q, niter, eps=1e-6, est_covs=True, est_var=True, ctx=ctx)
    assert q.iterations == niter
    for f in ("AHat", "BHat", "SigmaA", "SigmaB", "CA", "CB", "invCA", "invCB"):
        assert G.rel(getattr(q, f), jl(g["log_" + f], niter)) < TOL, f
    assert abs(q.sigma2 - g["log_sigma2"][niter]) < TOL * g["log_sigma2"][niter]
    assert G.rel(q.YHat, q.BHat @ q.AHat.T) < 1e-13


@pytest.mark.parametrize("niter", [1, 2, 10, 100])
def test_sparse_golden(G, ctx, niter):
    g = load_golden("sparse_test")
    Y, p = sparse_state_from_golden(g)
    q = G.to_gpu_params(p)
    d = G.vb.vbmf_sparse_(np.asfortranarray(Y), q, niter, eps=1e-6, full_cov=True, diag_var=False, est_cb=True, ctx=ctx,
                          keep_blocks=True)
    assert q.iterations == niter
    for f in ("AHat", "BHat", "SigmaA", "SigmaB"):
        assert G.rel(getattr(q, f), jl(g["log_" + f], niter)) < TOL, f
    for f in ("ATVecHat", "diagSigmaATVec", "CA", "beta", "CB", "delta"):
        assert G.rel(getattr(q, f), g["log_" + f][niter]) < TOL, f
    for f in ("sigmaHat", "zeta"):
        assert abs(getattr(q, f) - g["log_" + f][niter]) < TOL * abs(g["log_" + f][niter]), f
    assert G.rel(q.SigmaATVec_blocks, g["log_SigmaATVec"][niter]) < TOL
    assert d > 1e-6


# ------------------------------------------------------------------------------------------------ differential vs oracle
def _dense_case(L, M, H, seed, H1=0, labels=()):
    Y = synth(L, M, max(1, H // 2), seed)
    p = vo.vbmf_init(Y, H, ca=1.0, cb=1.0, sigma2=1.0, H1=H1, labels=labels, rng=np.random.default_rng(seed + 1))
    return Y, p


@pytest.mark.parametrize("L,M,H,H1,est", [(40, 150, 4, 0, (True, True)), (64, 500, 8, 2, (True, True)),
                                          (100, 257, 16, 0, (False, True)), (31, 77, 3, 1, (True, False)),
                                          (200, 1000, 64, 0, (True, True))])
def test_dense_vs_oracle(G, ctx, L, M, H, H1, est):
    labels = list(range(1, M + 1, 3)) if H1 else []
    Y, p = _dense_case(L, M, H, seed=L + M, H1=H1, labels=labels)
    Yf = np.asfortranarray(Y)
    # per-iteration parity (north_star): one GPU iteration from the oracle's state of every iteration
    po = copy.deepcopy(p)
    for i in range(12):
        q = G.to_gpu_params(po)
        vo.vbmf_run(Y, po, 1, eps=0.0, est_covs=est[0], est_var=est[1])
        G.vb.vbmf_(Yf, q, 1, eps=0.0, est_covs=est[0], est_var=est[1], ctx=ctx)
        G.compare(q, po, TOL)
    # free-running: agreement down to the trajectory's own sensitivity to 1e-14 perturbations
    for niter in (1, 3, 12):
        po = copy.deepcopy(p)
        _, it_o, d_o = vo.vbmf_run(Y, po, niter, eps=1e-9, est_covs=est[0], est_var=est[1])
        q = G.to_gpu_params(p)
        G.vb.vbmf_(Yf, q, niter, eps=1e-9, est_covs=est[0], est_var=est[1], ctx=ctx)
        assert q.iterations == it_o
        floor = G.sensitivity(lambda s: vo.vbmf_run(Y, s, niter, eps=1e-9, est_covs=est[0], est_var=est[1]), p, ["AHat", "BHat"])
        G.compare(q, po, max(TOL, 100 * floor))
        assert abs(q.d - d_o) <= max(1e-8, 1e4 * floor) * abs(d_o)


def test_dense_convergence_iteration(G, ctx):
    """Early exit: the device-side test must stop at the same iteration as the reference loop (spectral norm, Q1)."""
    Y, p = _dense_case(50, 300, 4, seed=11)
    for eps, norm in ((1e-3, "spectral"), (1e-4, "spectral"), (1e-3, "frobenius")):
        po = copy.deepcopy(p)
        _, it_o, d_o = vo.vbmf_run(Y, po, 500, eps=eps, est_covs=True, est_var=True, norm=norm)
        q = G.to_gpu_params(p)
        G.vb.vbmf_(np.asfortranarray(Y), q, 500, eps=eps, est_covs=True, est_var=True, norm=norm, ctx=ctx)
        assert it_o < 500
        assert q.iterations == it_o
        assert abs(q.d - d_o) <= 1e-7 * d_o
        # free-running for tens of iterations: bounded by the trajectory's own sensitivity over the compared fields
        G.compare(q, po, max(TOL, 100 * G.sensitivity(lambda s_: vo.vbmf_run(Y, s_, 500, eps=eps, est_covs=True, est_var=True, norm=norm), p,
                                                      G.FIELDS["dense"])))


SPARSE_CASES = [
    # L, M, H, H1, full_cov, diag_var, est_cb
    (30, 120, 4, 0, False, False, True),     # diagonal path incl. Q2 repeat(inner)
    (30, 120, 4, 0, True, False, True),
    (45, 200, 6, 2, False, False, False),    # labels / H1 mask, CB fixed
    (45, 200, 6, 2, True, True, True),       # heteroscedastic noise, full covariance
    (64, 333, 8, 0, False, True, True),      # heteroscedastic, diagonal (Q4)
    (80, 600, 32, 0, True, False, True),     # warp-per-matrix limit H = 32
    (40, 90, 40, 0, True, False, True),      # CTA-per-matrix path H > 32
]


@pytest.mark.parametrize("L,M,H,H1,full_cov,diag_var,est_cb", SPARSE_CASES)
def test_sparse_vs_oracle(G, ctx, L, M, H, H1, full_cov, diag_var, est_cb):
    Y = synth(L, M, max(1, H // 2), seed=L * M)
    labels = list(range(2, M + 1, 4)) if H1 else []
    p = vo.vbmf_sparse_init(Y, H, H1=H1, labels=labels, rng=np.random.default_rng(5))
    kw = dict(diag_var=diag_var, full_cov=full_cov, est_cb=est_cb)
    for niter in (1, 2, 8):
        po = copy.deepcopy(p)
        d_o, it_o = vo.vbmf_sparse_run(Y, po, niter, eps=1e-12, **kw)
        q = G.to_gpu_params(p)
        d = G.vb.vbmf_sparse_(np.asfortranarray(Y), q, niter, eps=1e-12, ctx=ctx, keep_blocks=full_cov, **kw)
        assert q.iterations == it_o
        floor = G.sensitivity(lambda s: vo.vbmf_sparse_run(Y, s, niter, eps=1e-12, **kw), p, ["AHat", "BHat"])
        G.compare(q, po, max(TOL, 100 * floor))
        assert abs(d - d_o) <= 1e-8 * abs(d_o)
        if full_cov:
            assert G.rel(q.SigmaATVec_blocks, po.SigmaATVec_blocks) < TOL
        assert G.rel(q.YHat, po.BHat @ po.AHat.T) < TOL


DUAL_CASES = [(30, 100, 4, 2, False, False, True, True), (30, 100, 4, 1, True, False, True, True),
              (50, 256, 8, 8, False, False, True, False), (50, 256, 8, 0, False, True, False, True),
              (36, 140, 6, 3, True, True, True, True)]


@pytest.mark.parametrize("L,M,H,H0,full_cov,diag_var,est_priors,est_cb", DUAL_CASES)
def test_dual_vs_oracle(G, ctx, L, M, H, H0, full_cov, diag_var, est_priors, est_cb):
    Y = synth(L, M, max(1, H // 2), seed=L + 3 * M)
    p = vo.vbmf_dual_init(Y, H, H0, rng=np.random.default_rng(9))
    for niter in (1, 2, 6):
        po = copy.deepcopy(p)
        d_o, it_o = vo.vbmf_dual_run(Y, po, niter, eps=1e-12, diag_var=diag_var, full_cov=full_cov, est_priors=est_priors, est_cb=est_cb)
        q = G.to_gpu_params(p)
        d = G.vb.vbmf_dual_(np.asfortranarray(Y), q, niter, eps=1e-12, diag_var=diag_var, full_cov=full_cov,
                            est_priors=est_priors, est_cb=est_cb, ctx=ctx)
        assert q.iterations == it_o
        floor = G.sensitivity(lambda s: vo.vbmf_dual_run(Y, s, niter, eps=1e-12, diag_var=diag_var, full_cov=full_cov,
                                                         est_priors=est_priors, est_cb=est_cb), p, ["AHat", "BHat"])
        G.compare(q, po, max(TOL, 100 * floor))
        assert abs(d - d_o) <= 1e-8 * abs(d_o)


def _teacher_forced(G, ctx, Y, p, run_o, run_g, iters=10):
    """north_star's bar as written: every iteration starts from the ORACLE's state of that iteration, one GPU iteration, every
    field <= 1e-10 (the free-running loops above additionally bound the drift by the trajectory's own sensitivity)."""
    po = copy.deepcopy(p)
    for _ in range(iters):
        q = G.to_gpu_params(po)
        old = po.BHat.copy()
        d_o = run_o(po)
        d = run_g(q)
        assert q.iterations == 1
        G.compare(q, po, TOL)
        G.record("delta_abs", abs(d - d_o))
        assert abs(d - d_o) <= 1e-9 + 1e-8 * abs(d_o)


@pytest.mark.parametrize("L,M,H,H1,full_cov,diag_var,est_cb", SPARSE_CASES)
def test_sparse_teacher_forced(G, ctx, L, M, H, H1, full_cov, diag_var, est_cb):
    Y = synth(L, M, max(1, H // 2), seed=L * M)
    labels = list(range(2, M + 1, 4)) if H1 else []
    p = vo.vbmf_sparse_init(Y, H, H1=H1, labels=labels, rng=np.random.default_rng(5))
    kw = dict(diag_var=diag_var, full_cov=full_cov, est_cb=est_cb)
    Yf = np.asfortranarray(Y)
    _teacher_forced(G, ctx, Y, p, lambda s: vo.vbmf_sparse_run(Y, s, 1, eps=0.0, **kw)[0],
                    lambda q: G.vb.vbmf_sparse_(Yf, q, 1, eps=0.0, ctx=ctx, yhat=False, **kw))


@pytest.mark.parametrize("L,M,H,H0,full_cov,diag_var,est_priors,est_cb", DUAL_CASES)
def test_dual_teacher_forced(G, ctx, L, M, H, H0, full_cov, diag_var, est_priors, est_cb):
    Y = synth(L, M, max(1, H // 2), seed=L + 3 * M)
    p = vo.vbmf_dual_init(Y, H, H0, rng=np.random.default_rng(9))
    kw = dict(diag_var=diag_var, full_cov=full_cov, est_priors=est_priors, est_cb=est_cb)
    Yf = np.asfortranarray(Y)
    _teacher_forced(G, ctx, Y, p, lambda s: vo.vbmf_dual_run(Y, s, 1, eps=0.0, **kw)[0],
                    lambda q: G.vb.vbmf_dual_(Yf, q, 1, eps=0.0, ctx=ctx, yhat=False, **kw))


def test_dual_H_lt_H0_errors(G, ctx):
    Y = synth(10, 20, 2, 0)
    with pytest.raises(G.vb.VBMFError, match="H must be at least H0"):
        G.vb.vbmf_dual_init(Y, 2, 3)


# ------------------------------------------------------------------------------------------------ lower bound
@pytest.mark.parametrize("kind", ["sparse", "dual"])
@pytest.mark.parametrize("full_cov", [False, True])
def test_lower_bound(G, ctx, kind, full_cov):
    L, M, H = 12, 60, 4
    Y = synth(L, M, 2, seed=21)
    if kind == "sparse":
        p = vo.vbmf_sparse_init(Y, H, rng=np.random.default_rng(2))
        vo.vbmf_sparse_run(Y, p, 15, eps=0.0, full_cov=full_cov)
        lb_o, lbt_o = vo.sparse_lowerBound(Y, p), vo.sparse_lowerBoundTrimmed(Y, p, 1e-1)
    else:
        p = vo.vbmf_dual_init(Y, H, 2, rng=np.random.default_rng(2))
        vo.vbmf_dual_run(Y, p, 15, eps=0.0, full_cov=full_cov)
        lb_o, lbt_o = vo.dual_lowerBound(Y, p), vo.dual_lowerBoundTrimmed(Y, p, 1e-1)
    q = G.to_gpu_params(p)
    Yf = np.asfortranarray(Y)
    lb = G.vb.lowerBound(Yf, q, ctx=ctx)
    lbt = G.vb.lowerBoundTrimmed(Yf, q, 1e-1, ctx=ctx)
    assert abs(lb - lb_o) <= TOL * abs(lb_o), (lb, lb_o)
    assert abs(lbt - lbt_o) <= TOL * abs(lbt_o), (lbt, lbt_o)


def test_lower_bound_golden_state(G, ctx):
    """ELBO of the reference's own logged final state (sparse fixture, iteration 100): oracle vs GPU."""
    g = load_golden("sparse_test")
    Y, p = sparse_state_from_golden(g, 100)
    lb_o = vo.sparse_lowerBound(Y, p)
    lb = G.vb.lowerBound(np.asfortranarray(Y), G.to_gpu_params(p), ctx=ctx)
    assert abs(lb - lb_o) <= TOL * abs(lb_o)


# ------------------------------------------------------------------------------------------------ step-level API (MIL call pattern)
@pytest.mark.parametrize("kind", ["dense", "sparse", "dual"])
def test_vbls_steps(G, ctx, kind):
    """examples/mil_util.jl:179-203: updateA!, updateCA!, updateSigma*! with BHat fixed, called one by one."""
    L, M, H = 20, 37, 4
    Y = synth(L, M, 2, seed=33)
    rng = np.random.default_rng(4)
    if kind == "dense":
        p = vo.vbmf_init(Y, H, rng=rng)
        kw = {}
    elif kind == "sparse":
        p = vo.vbmf_sparse_init(Y, H, rng=rng)
        kw = {"full_cov": True}
    else:
        p = vo.vbmf_dual_init(Y, H, 2, rng=rng)
        kw = {"full_cov": True}
    q = G.to_gpu_params(p)
    vo.vbls(Y, p, 5, **kw)
    A = G.vb.vbls_(np.asfortranarray(Y), q, 5, ctx=ctx, **kw)
    assert G.rel(A, p.AHat) < TOL
    G.compare(q, p, TOL)
    assert G.rel(q.YHat, p.YHat) < TOL


def test_single_steps_sparse(G, ctx):
    L, M, H = 25, 80, 6
    Y = synth(L, M, 3, seed=44)
    Yf = np.asfortranarray(Y)
    p = vo.vbmf_sparse_init(Y, H, rng=np.random.default_rng(6))
    q = G.to_gpu_params(p)
    vo.sparse_updateA(Y, p); G.vb.updateA_(Yf, q, ctx=ctx); G.compare(q, p, TOL)
    vo.sparse_updateB(Y, p); G.vb.updateB_(Yf, q, ctx=ctx); G.compare(q, p, TOL)
    vo.sparse_updateCA(p); G.vb.updateCA_(q, ctx=ctx); G.compare(q, p, TOL)
    vo.sparse_updateCB(p); G.vb.updateCB_(q, ctx=ctx); G.compare(q, p, TOL)
    vo.sparse_updateSigma(Y, p); G.vb.updateSigma_(Yf, q, ctx=ctx); G.compare(q, p, TOL)


def test_no_y_attached_errors(G):
    c = G.vb.Context(device=0)
    p = G.vb.vbmf_init(np.zeros((4, 5)), 2)
    with pytest.raises(G.vb.VBMFError):
        G.vb.Solver(c, p)
    c.close()


# ------------------------------------------------------------------------------------------------ batched small problems (MIL pattern)
@pytest.mark.parametrize("kind,full_cov,H", [("dual", True, 20), ("dual", False, 20), ("sparse", True, 20), ("sparse", False, 7),
                                             ("sparse", True, 32), ("dual", True, 4), ("sparse", True, 7), ("dual", True, 13),
                                             ("trial", True, 12), ("trial", False, 9), ("dense", False, 20), ("dense", False, 5)])
def test_vbls_batched(G, ctx, kind, full_cov, H):
    """examples/mil_util.jl:504-511: vbls! on every bag x class model; one CTA per problem must equal the oracle's vbls."""
    rng = np.random.default_rng(17)
    L, nprob, niter = 38, 40, 20
    Ys, ps, qs = [], [], []
    for b in range(nprob):
        M = int(rng.integers(2, 41)) if b else 1          # includes the M = 1 edge (Q2 degenerates)
        Y = 10.0 * synth(L, M, 3, seed=100 + b)           # scale = 10 as in examples/mil_data.jl:19
        if kind == "sparse":
            p = vo.vbmf_sparse_init(Y, H, rng=rng)
        elif kind == "dense":
            p = vo.vbmf_init(Y, H, sigma2=0.9, rng=rng)   # copy_vbmf_params: vbmf_init(Y, H, sigma2 = old.sigma2), mil_util.jl:216
        elif kind == "trial":                             # copy_vbmf_params uses M0 = M (:254); also exercise a real row split
            p = vo.vbmf_trial_init(Y, H, H - 2, M if b % 2 else M // 2, rng=rng)
            p.alpha01, p.beta01, p.alpha02, p.beta02, p.alpha03, p.beta03 = 0.3, 0.2, 0.7, 0.4, 1e-10, 1e-10
        else:
            p = vo.vbmf_dual_init(Y, H, H - 1, rng=rng)   # H1 = 1 as in examples/mil_data.jl:23
        p.SigmaB = np.diag(rng.uniform(1e-3, 1e-2, H))    # a trained model carries a non-trivial SigmaB
        if kind != "dense":
            p.sigmaHat = 0.7
        Ys.append(np.asfortranarray(Y)); ps.append(p); qs.append(G.to_gpu_params(p))
    G.vb.vbls_batched_(Ys, qs, niter, full_cov=full_cov, ctx=ctx)
    fields = ["AHat", "ATVecHat", "diagSigmaATVec", "SigmaA", "CA", "beta", "sigmaHat", "zeta"]
    if kind == "dual":
        fields += ["A0Hat", "A1Hat", "CA0", "CA1", "beta0", "beta1", "alpha0", "alpha1"]
    if kind == "trial":
        fields += ["A1Hat", "A2Hat", "A3Hat", "CA1", "CA2", "CA3", "beta1", "beta2", "beta3", "alpha1", "alpha2", "alpha3"]
    if kind == "dense":
        fields = ["AHat", "SigmaA", "CA", "invCA", "sigma2"]
    for Y, p, q in zip(Ys, ps, qs):
        vo.vbls(Y, p, niter, full_cov=full_cov)
        G.compare(q, p, TOL, fields)
        assert G.rel(q.YHat, p.YHat) < TOL


# ------------------------------------------------------------------------------------------------ preprocess / scaleY (N3)
def test_preprocess(G, ctx):
    """src/util.jl:36-87: row standardisation with the 1e-15 / 1e-8 guards, removal of near-constant rows, times lambda."""
    rng = np.random.default_rng(8)
    L, M = 57, 333
    Y = rng.standard_normal((L, M)) * rng.uniform(0.1, 50.0, (L, 1)) + rng.uniform(-5, 5, (L, 1))
    Y[7, :] = 3.25            # constant row -> variance 0 -> scaled row is all zeros -> dropped
    Y[20, :] = 0.0
    Y[33, :] = 1.0 + 1e-12 * rng.standard_normal(M)   # |Y - mu| <= 1e-8 -> zeroed -> dropped
    ref, used = vo.preprocess(Y, 10.0)
    out = G.vb.preprocess(np.asfortranarray(Y), 10.0, ctx=ctx)
    assert out.shape == ref.shape == (L - 3, M)
    assert G.rel(out, ref) < 1e-12
    # and the resident matrix is the processed one
    assert abs(ctx.trYTY() - float(np.sum(ref * ref))) <= 1e-11 * float(np.sum(ref * ref))
    rows = None
    ctx.attach(np.asfortranarray(Y), force=True)
    rows = ctx.preprocess(10.0)
    assert np.array_equal(rows, used)
    # no row dropped: in-place path
    Y2 = rng.standard_normal((16, 40))
    ref2, used2 = vo.preprocess(Y2, 2.0)
    assert G.rel(G.vb.preprocess(np.asfortranarray(Y2), 2.0, ctx=ctx), ref2) < 1e-12 and used2.size == 16


# ------------------------------------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("L,M,H", [(1, 5, 1), (7, 1, 2), (3, 2, 3), (5, 9, 1), (2, 2, 2)])
def test_degenerate_shapes(G, ctx, L, M, H):
    """Smallest shapes the reference accepts: single row / single column (Q2's repeat(inner = 0)) / rank one."""
    Y = synth(L, M, 1, seed=L * 10 + M)
    Yf = np.asfortranarray(Y)
    p = vo.vbmf_init(Y, H, rng=np.random.default_rng(1))
    q = G.to_gpu_params(p)
    vo.vbmf_run(Y, p, 4, eps=0.0, est_covs=True, est_var=True)
    G.vb.vbmf_(Yf, q, 4, eps=0.0, est_covs=True, est_var=True, ctx=ctx)
    G.compare(q, p, TOL)
    for full_cov in (False, True):
        ps = vo.vbmf_sparse_init(Y, H, rng=np.random.default_rng(2))
        qs = G.to_gpu_params(ps)
        vo.vbmf_sparse_run(Y, ps, 3, eps=0.0, full_cov=full_cov)
        G.vb.vbmf_sparse_(Yf, qs, 3, eps=0.0, full_cov=full_cov, ctx=ctx)
        G.compare(qs, ps, TOL)


def test_zero_iterations_and_eps_edge(G, ctx):
    """niter = 0 leaves the state untouched (d = eps + 1); eps = Inf never enters the loop (eps + 1 > eps is false)."""
    Y = synth(12, 30, 2, seed=3)
    p = vo.vbmf_init(Y, 3, rng=np.random.default_rng(1))
    for niter, eps in ((0, 1e-6), (10, float("inf"))):
        q = G.to_gpu_params(p)
        G.vb.vbmf_(np.asfortranarray(Y), q, niter, eps=eps, est_covs=True, est_var=True, ctx=ctx)
        assert q.iterations == 0
        assert np.array_equal(q.AHat, p.AHat) and np.array_equal(q.BHat, p.BHat) and q.sigma2 == p.sigma2
        assert G.rel(q.YHat, p.BHat @ p.AHat.T) < 1e-14      # updateYHat! still runs after the loop


def test_dual_prior_steps(G, ctx):
    """updateAlpha00!/01!, updateBeta00!/01! called one by one (src/vbmf_dual.jl:393-434) after a few iterations."""
    Y = synth(20, 64, 2, seed=5)
    p = vo.vbmf_dual_init(Y, 4, 2, rng=np.random.default_rng(3))
    vo.vbmf_dual_run(Y, p, 3, eps=0.0, est_priors=False)
    q = G.to_gpu_params(p)
    for fo, fg in ((vo.dual_updateAlpha00, G.vb.updateAlpha00_), (vo.dual_updateAlpha01, G.vb.updateAlpha01_),
                   (vo.dual_updateBeta00, G.vb.updateBeta00_), (vo.dual_updateBeta01, G.vb.updateBeta01_)):
        fo(p)
        fg(q, ctx=ctx)
        G.compare(q, p, TOL, ["alpha00", "alpha01", "beta00", "beta01", "CA", "beta"])


def test_error_paths(G, ctx):
    Y = np.asfortranarray(synth(8, 10, 2, seed=1))
    with pytest.raises(G.vb.VBMFError, match="outside the supported range"):
        G.vb.vbmf_(Y, G.vb.vbmf_init(Y, 129), 1, ctx=ctx)
    p = G.vb.vbmf_init(Y, 2)
    p.M = 11
    with pytest.raises((G.vb.VBMFError, ValueError)):
        G.vb.vbmf_(Y, p, 1, ctx=ctx)
    with pytest.raises(G.vb.VBMFError, match="no lowerBound"):
        G.vb.lowerBound(Y, G.vb.vbmf_init(Y, 2), ctx=ctx)
    # a non positive definite precision (negative prior covariance) must surface as NaN + ended loop, not as garbage
    q = G.vb.vbmf_init(Y, 2, rng=np.random.default_rng(0))
    q.invCA = np.asfortranarray(-1e6 * np.eye(2))
    G.vb.vbmf_(Y, q, 5, eps=0.0, ctx=ctx)
    assert q.iterations == 1 and np.isnan(q.d) and np.isnan(q.AHat).all()


# ------------------------------------------------------------------------------------------------ trajectory logging (N4) vs the reference's own log
def test_logged_trajectory_matches_reference_log(G, ctx, tmp_path):
    """vbmf_ with logdir: every logged field of every one of the 100 iterations against examples/data/vbmf_test/log.jld
    (create_log / update_log! / save_log, src/data_manip.jl:6-66), then load_log + extract_params! round trip."""
    g = load_golden("vbmf_test")
    Y, p = dense_state_from_golden(g)
    q = G.to_gpu_params(p)
    G.vb.vbmf_(np.asfortranarray(Y), q, 100, eps=1e-6, est_covs=True, est_var=True, ctx=ctx, logdir=str(tmp_path), desc="run")
    assert q.iterations == 100
    log, Yl = G.vb.load_log(str(tmp_path / "run"))
    assert np.array_equal(Yl, Y) and log["sigma2"].shape == (101,) and log["AHat"].shape == (20, 2, 101)
    assert G.rel(log["sigma2"], g["log_sigma2"]) < TOL
    for f in ("AHat", "BHat", "SigmaA", "SigmaB", "CA", "CB"):
        ref = np.stack([jl(g["log_" + f], t) for t in range(101)], axis=-1)
        assert G.rel(log[f], ref) < TOL, f
    r = G.to_gpu_params(p)
    G.vb.extract_params_(log, r, 37)
    assert G.rel(r.BHat, jl(g["log_BHat"], 37)) < TOL and abs(r.sigma2 - g["log_sigma2"][37]) < TOL



# ------------------------------------------------------------------------------------------------ vbmf_trial (N2)
TRIAL_CASES = [(30, 100, 4, 2, 40, False, False, True), (30, 100, 4, 1, 70, True, False, True), (36, 140, 6, 3, 50, True, True, True),
               (25, 64, 5, 2, 0, False, False, True), (25, 64, 5, 2, 64, False, False, False)]


@pytest.mark.parametrize("L,M,H,H0,M0,full_cov,diag_var,est_priors", TRIAL_CASES)
def test_trial_vs_oracle(G, ctx, L, M, H, H0, M0, full_cov, diag_var, est_priors):
    """src/vbmf_trial.jl: three ARD groups (column split H0, row split M0) with learned hyper-priors; incl. the empty-group edges."""
    Y = synth(L, M, max(1, H // 2), seed=L + 5 * M)
    p = vo.vbmf_trial_init(Y, H, H0, M0, rng=np.random.default_rng(13))
    for niter in (1, 2, 6):
        po = copy.deepcopy(p)
        d_o, it_o = vo.vbmf_trial_run(Y, po, niter, eps=1e-12, diag_var=diag_var, full_cov=full_cov, est_priors=est_priors)
        q = G.to_gpu_params(p)
        d = G.vb.vbmf_trial_(np.asfortranarray(Y), q, niter, eps=1e-12, diag_var=diag_var, full_cov=full_cov, est_priors=est_priors, ctx=ctx)
        assert q.iterations == it_o
        floor = G.sensitivity(lambda s: vo.vbmf_trial_run(Y, s, niter, eps=1e-12, diag_var=diag_var, full_cov=full_cov,
                                                          est_priors=est_priors), p, ["AHat", "BHat"])
        G.compare(q, po, max(TOL, 100 * floor))
        assert abs(d - d_o) <= 1e-8 * abs(d_o) + 1e-9
    if not diag_var:
        lb_o, lbt_o = vo.trial_lowerBound(Y, po), vo.trial_lowerBoundTrimmed(Y, po, 1e-1)
        lb, lbt = G.vb.lowerBound(np.asfortranarray(Y), q, ctx=ctx), G.vb.lowerBoundTrimmed(np.asfortranarray(Y), q, 1e-1, ctx=ctx)
        if np.isfinite(lb_o):
            assert abs(lb - lb_o) <= 1e-9 * abs(lb_o) and abs(lbt - lbt_o) <= 1e-9 * abs(lbt_o), (lb, lb_o, lbt, lbt_o)
        else:
            assert np.isnan(lb) == np.isnan(lb_o)


@pytest.mark.parametrize("L,M,H,H0,M0,full_cov,diag_var,est_priors", TRIAL_CASES)
def test_trial_teacher_forced(G, ctx, L, M, H, H0, M0, full_cov, diag_var, est_priors):
    Y = synth(L, M, max(1, H // 2), seed=L + 5 * M)
    p = vo.vbmf_trial_init(Y, H, H0, M0, rng=np.random.default_rng(13))
    kw = dict(diag_var=diag_var, full_cov=full_cov, est_priors=est_priors)
    Yf = np.asfortranarray(Y)
    _teacher_forced(G, ctx, Y, p, lambda s: vo.vbmf_trial_run(Y, s, 1, eps=0.0, **kw)[0],
                    lambda q: G.vb.vbmf_trial_(Yf, q, 1, eps=0.0, ctx=ctx, yhat=False, **kw))


# ------------------------------------------------------------------------------------------------ wide ranks (BN = 128 tiles, padded epilogues)
@pytest.mark.parametrize("kind,L,M,H", [("dense", 300, 700, 100), ("dense", 1024, 4096, 128), ("dual", 2048, 8192, 128),
                                        ("sparse", 200, 900, 77), ("sparse_full", 150, 260, 48)])
def test_wide_rank(G, ctx, kind, L, M, H):
    """H up to 128 (config 5's rank), H not a multiple of 8 / 16 (padded DMMA epilogues), odd H (SIMT K2), CTA-per-matrix K4."""
    Y = synth(L, M, max(2, H // 4), seed=H)
    Yf = np.asfortranarray(Y)
    rng = np.random.default_rng(H)
    if kind == "dense":
        p = vo.vbmf_init(Y, H, rng=rng)
        q = G.to_gpu_params(p)
        vo.vbmf_run(Y, p, 2, eps=0.0, est_covs=True, est_var=True)
        G.vb.vbmf_(Yf, q, 2, eps=0.0, est_covs=True, est_var=True, ctx=ctx, yhat=False)
    elif kind == "dual":
        p = vo.vbmf_dual_init(Y, H, H // 2, rng=rng)
        q = G.to_gpu_params(p)
        vo.vbmf_dual_run(Y, p, 2, eps=0.0)
        G.vb.vbmf_dual_(Yf, q, 2, eps=0.0, ctx=ctx, yhat=False)
    else:
        full = kind == "sparse_full"
        p = vo.vbmf_sparse_init(Y, H, rng=rng)
        q = G.to_gpu_params(p)
        vo.vbmf_sparse_run(Y, p, 2, eps=0.0, full_cov=full)
        G.vb.vbmf_sparse_(Yf, q, 2, eps=0.0, full_cov=full, ctx=ctx, yhat=False)
    G.compare(q, p, TOL)


def test_bit_reproducible_runs(G, ctx):
    """Fixed-order reductions everywhere (no atomics), also with the side-stream tail and the CUDA-graph replay: two runs
    from the same state give bit-identical results."""
    Y = synth(300, 5000, 8, seed=77)
    Yf = np.asfortranarray(Y)
    p = vo.vbmf_sparse_init(Y, 16, rng=np.random.default_rng(1))
    outs = []
    for _ in range(2):
        q = G.to_gpu_params(p)
        G.vb.vbmf_sparse_(Yf, q, 12, eps=0.0, full_cov=True, ctx=ctx, yhat=False)
        outs.append(q)
    for f in ("AHat", "BHat", "CA", "SigmaA", "SigmaB", "sigmaHat", "zeta"):
        assert np.array_equal(np.asarray(getattr(outs[0], f)), np.asarray(getattr(outs[1], f))), f


def test_batched_staging_reuse_and_growth(G, ctx):
    """The batched path keeps a grow-only device arena + pinned mirror in the context: a small batch, a larger one (regrowth)
    and the small one again must all give the oracle's answer (no stale bytes from the earlier layout)."""
    rng = np.random.default_rng(3)
    L, H, niter = 12, 5, 6

    def make(n, seed):
        Ys, ps = [], []
        for b in range(n):
            Y = synth(L, int(rng.integers(2, 30)), 2, seed=seed + b)
            p = vo.vbmf_sparse_init(Y, H, rng=rng)
            p.SigmaB = np.diag(rng.uniform(1e-3, 1e-2, H))
            Ys.append(np.asfortranarray(Y)); ps.append(p)
        return Ys, ps
    for n, seed in ((3, 10), (60, 200), (3, 10), (7, 900)):
        Ys, ps = make(n, seed)
        qs = [G.to_gpu_params(p) for p in ps]
        batch = G.vb.BatchedVbls(Ys, qs, ctx=ctx, yhat=True, keep_blocks=True)
        batch.run(niter, full_cov=True)
        batch.readback()
        for Y, p, q in zip(Ys, ps, qs):
            vo.vbls(Y, p, niter, full_cov=True)
            G.compare(q, p, TOL, ["AHat", "diagSigmaATVec", "SigmaA", "CA", "beta", "sigmaHat"])
            assert G.rel(q.SigmaATVec_blocks, p.SigmaATVec_blocks) < TOL


def test_no_device_memory_leak(G):
    """Contexts, solvers, graphs, side streams and staging buffers are all released: 40 create/run/destroy cycles of every
    solver kind leave the device's free memory where it was."""
    import torch
    Y = synth(200, 3000, 4, seed=5)
    Yf = np.asfortranarray(Y)

    def cycle():
        c = G.vb.Context(device=0)
        p = vo.vbmf_init(Y, 8, rng=np.random.default_rng(0))
        G.vb.vbmf_(Yf, G.to_gpu_params(p), 3, eps=0.0, est_covs=True, est_var=True, ctx=c, yhat=False)
        s = vo.vbmf_sparse_init(Y, 8, rng=np.random.default_rng(0))
        G.vb.vbmf_sparse_(Yf, G.to_gpu_params(s), 3, eps=0.0, full_cov=True, ctx=c, yhat=False)
        d = vo.vbmf_dual_init(Y, 8, 4, rng=np.random.default_rng(0))
        G.vb.vbmf_dual_(Yf, G.to_gpu_params(d), 3, eps=0.0, ctx=c, yhat=False)
        G.vb.vbls_batched_([Yf[:, :20]], [G.to_gpu_params(vo.vbmf_sparse_init(Y[:, :20], 8, rng=np.random.default_rng(0)))], 2,
                           full_cov=True, ctx=c)
        c.close()
    for _ in range(3):
        cycle()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info(0)[0]
    for _ in range(40):
        cycle()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info(0)[0]
    assert free0 - free1 < 8 << 20, "device memory shrank by %.1f MB over 40 cycles" % ((free0 - free1) / 2**20)


def _fuzz_shapes():
    rng = np.random.default_rng(20261018)
    Hs = [1, 3, 5, 7, 9, 11, 13, 17, 23, 31, 33, 37, 47, 63, 65, 67, 97, 127]
    out = []
    for H in Hs:
        L = int(rng.integers(1, 160)) * 2 + 1                # odd
        M = int(rng.integers(H + 1, 900))
        out.append((L, M, H))
    return out


@pytest.mark.parametrize("L,M,H", _fuzz_shapes())
def test_odd_shape_fuzz(G, ctx, L, M, H):
    """Odd / prime L, M, H through every kernel family (padding, alignment and tail handling): two teacher-forced iterations
    of dense, sparse (diagonal and full covariance, heteroscedastic) and dual against the oracle."""
    Y = synth(L, M, max(1, min(H, L) // 2), seed=L + 7 * M + H)
    Yf = np.asfortranarray(Y)
    p = vo.vbmf_init(Y, H, rng=np.random.default_rng(H))
    q = G.to_gpu_params(p)
    vo.vbmf_run(Y, p, 2, eps=0.0, est_covs=True, est_var=True)
    G.vb.vbmf_(Yf, q, 2, eps=0.0, est_covs=True, est_var=True, ctx=ctx, yhat=False)
    G.compare(q, p, TOL, ["AHat", "BHat", "SigmaA", "SigmaB", "CA", "CB", "sigma2"])
    for full_cov, diag_var in ((False, False), (True, False), (True, True), (False, True)):
        ps = vo.vbmf_sparse_init(Y, H, rng=np.random.default_rng(H + 1))
        qs = G.to_gpu_params(ps)
        vo.vbmf_sparse_run(Y, ps, 2, eps=0.0, full_cov=full_cov, diag_var=diag_var)
        G.vb.vbmf_sparse_(Yf, qs, 2, eps=0.0, full_cov=full_cov, diag_var=diag_var, ctx=ctx, yhat=False)
        G.compare(qs, ps, TOL, ["AHat", "BHat", "SigmaA", "SigmaB", "CA", "beta", "CB", "diagSigmaATVec", "sigmaHat", "sigmaVecHat"])
    H0 = H // 2
    pd = vo.vbmf_dual_init(Y, H, H0, rng=np.random.default_rng(H + 2))
    qd = G.to_gpu_params(pd)
    vo.vbmf_dual_run(Y, pd, 2, eps=0.0, full_cov=(H % 4 == 1))
    G.vb.vbmf_dual_(Yf, qd, 2, eps=0.0, full_cov=(H % 4 == 1), ctx=ctx, yhat=False)
    G.compare(qd, pd, TOL, ["AHat", "BHat", "SigmaA", "SigmaB", "CA", "CA0", "CA1", "beta", "CB", "sigmaHat", "alpha00", "alpha01", "beta00", "beta01"])


# ------------------------------------------------------------------------------------------------ MIL callers (N1)
def _oracle_copy_dual(Y, old):
    p = vo.vbmf_dual_init(Y, old.H, old.H0, gamma0=old.gamma0, delta0=old.delta0, eta0=old.eta0, zeta0=old.zeta0,
                          rng=np.random.default_rng(0))
    p.BHat, p.SigmaB, p.CB, p.gamma, p.delta = old.BHat.copy(), old.SigmaB.copy(), old.CB.copy(), old.gamma, old.delta.copy()
    p.alpha00, p.beta00, p.alpha01, p.beta01 = old.alpha00, old.beta00, old.alpha01, old.beta01
    return p


def test_mil_classify_dual_matches_reference_procedure(G, ctx):
    """examples/mil_util.jl: copy_vbmf_params (:237-251) + vbls!(..., 20, full_cov = true) on both class models + the
    spectral-norm decision rule (:513-533), for a whole test set in batched launches, against the same procedure spelled
    out with the oracle one bag at a time."""
    rng = np.random.default_rng(99)
    L, H, H0 = 24, 6, 5
    models = []
    for c in range(2):                                     # two "trained" class models
        Yt = 10.0 * synth(L, 60, 3, seed=500 + c)
        p = vo.vbmf_dual_init(Yt, H, H0, rng=rng)
        vo.vbmf_dual_run(Yt, p, 15, eps=0.0, full_cov=True)
        models.append(p)
    gm = [G.to_gpu_params(m) for m in models]
    bags = [np.asfortranarray(10.0 * synth(L, int(rng.integers(3, 25)), 3, seed=(500 if b % 2 == 0 else 501) + 7 * b)) for b in range(30)]
    labels, e0, e1 = G.vb.mil.classify_bags(gm[0], gm[1], bags, class_alg="dual", ctx=ctx)
    for b, Y in enumerate(bags):
        errs = []
        for m in models:
            p = _oracle_copy_dual(Y, m)
            vo.vbls(Y, p, 20, full_cov=True)
            errs.append(np.linalg.norm(p.YHat - Y, 2) / (Y.shape[0] * Y.shape[1]))
        assert abs(e0[b] - errs[0]) <= 1e-9 * errs[0] and abs(e1[b] - errs[1]) <= 1e-9 * errs[1]
        assert labels[b] == (0 if errs[0] < errs[1] else 1)
    lab1, a0, a1 = G.vb.mil.classify(gm[0], gm[1], bags[3], class_alg="dual", ctx=ctx)
    assert lab1 == labels[3] and a0 == e0[3] and a1 == e1[3]
    mer, eer, fp, fn, n0, n1 = G.vb.mil.test_classification(gm[0], gm[1], bags, labels, class_alg="dual", ctx=ctx)
    assert (mer, fp, fn) == (0.0, 0, 0) and n0 + n1 == len(bags)


def test_mil_train_dual_restarts(G, ctx):
    """train_dual (:327-385): column subsampling to floor(3200/H), vbmf_dual! with full_cov, and a usable (non-collapsed) model."""
    rng = np.random.default_rng(5)
    Y0, Y1 = 10.0 * synth(20, 900, 3, seed=1), 10.0 * synth(20, 50, 3, seed=2)
    p0, p1 = G.vb.mil.train_dual(Y0, Y1, 8, 7, 10, eps=1e-4, ctx=ctx, rng=rng)
    assert p0.M == 400 and p1.M == 50 and p0.H0 == 7
    for p in (p0, p1):
        assert np.linalg.norm(p.AHat, 2) + np.linalg.norm(p.BHat, 2) >= 1e-2 and np.all(np.isfinite(p.BHat))


def _random_cases():
    """Random (kind, shape, flags) cases; VBMF_FUZZ_N widens the sweep for a soak run (default keeps the suite fast)."""
    import os
    n = int(os.environ.get("VBMF_FUZZ_N", "8"))
    rng = np.random.default_rng(int(os.environ.get("VBMF_FUZZ_SEED", "424242")))
    out = []
    for c in range(n):
        kind = ["dense", "sparse", "dual", "trial"][int(rng.integers(0, 4))]
        H = int(rng.choice([1, 2, 3, 4, 6, 8, 12, 16, 20, 24, 31, 32, 33, 40, 48, 64, 65, 80, 96, 128]))
        L = int(rng.integers(1, 400))
        M = int(rng.integers(max(2, H // 4), 2500))
        full_cov = bool(rng.integers(0, 2)) and (H <= 64 or M <= 600)
        diag_var = bool(rng.integers(0, 2))
        out.append((c, kind, L, M, H, full_cov, diag_var))
    return out


@pytest.mark.parametrize("case,kind,L,M,H,full_cov,diag_var", _random_cases())
def test_random_fuzz(G, ctx, case, kind, L, M, H, full_cov, diag_var):
    """Random kinds / shapes / flags, one teacher-forced iteration twice (the second from the GPU's own state) vs the oracle."""
    rng = np.random.default_rng(1000 + case)
    Y = synth(L, M, max(1, min(H, L, M) // 2), seed=31 * case + 7)
    Yf = np.asfortranarray(Y)
    if kind == "dense":
        H1 = int(rng.integers(0, H + 1)) if case % 2 else 0
        labels = sorted(set(int(x) for x in rng.integers(1, M + 1, size=min(M, 5)))) if H1 else []
        p = vo.vbmf_init(Y, H, H1=H1, labels=labels, rng=rng)
        run_o = lambda q_: vo.vbmf_run(Y, q_, 1, eps=0.0, est_covs=True, est_var=True)
        run_g = lambda q_: G.vb.vbmf_(Yf, q_, 1, eps=0.0, est_covs=True, est_var=True, ctx=ctx, yhat=False)
        fields = ["AHat", "BHat", "SigmaA", "SigmaB", "CA", "CB", "sigma2"]
    elif kind == "sparse":
        H1 = int(rng.integers(0, H + 1)) if case % 2 else 0
        labels = sorted(set(int(x) for x in rng.integers(1, M + 1, size=min(M, 5)))) if H1 else []
        p = vo.vbmf_sparse_init(Y, H, H1=H1, labels=labels, rng=rng)
        run_o = lambda q_: vo.vbmf_sparse_run(Y, q_, 1, eps=0.0, full_cov=full_cov, diag_var=diag_var)
        run_g = lambda q_: G.vb.vbmf_sparse_(Yf, q_, 1, eps=0.0, full_cov=full_cov, diag_var=diag_var, ctx=ctx, yhat=False)
        fields = ["AHat", "BHat", "SigmaA", "SigmaB", "CA", "beta", "CB", "diagSigmaATVec", "sigmaHat", "sigmaVecHat"]
    elif kind == "dual":
        p = vo.vbmf_dual_init(Y, H, int(rng.integers(0, H + 1)), rng=rng)
        run_o = lambda q_: vo.vbmf_dual_run(Y, q_, 1, eps=0.0, full_cov=full_cov, diag_var=diag_var)
        run_g = lambda q_: G.vb.vbmf_dual_(Yf, q_, 1, eps=0.0, full_cov=full_cov, diag_var=diag_var, ctx=ctx, yhat=False)
        fields = ["AHat", "BHat", "SigmaA", "SigmaB", "CA", "beta", "CB", "sigmaHat", "sigmaVecHat", "alpha00", "alpha01", "beta00", "beta01"]
    else:
        p = vo.vbmf_trial_init(Y, H, int(rng.integers(0, H + 1)), int(rng.integers(0, M + 1)), rng=rng)
        run_o = lambda q_: vo.vbmf_trial_run(Y, q_, 1, eps=0.0, full_cov=full_cov, diag_var=diag_var)
        run_g = lambda q_: G.vb.vbmf_trial_(Yf, q_, 1, eps=0.0, full_cov=full_cov, diag_var=diag_var, ctx=ctx, yhat=False)
        fields = ["AHat", "BHat", "SigmaA", "SigmaB", "CA", "beta", "CB", "sigmaHat", "sigmaVecHat", "alpha01", "alpha02", "alpha03"]
    q = G.to_gpu_params(p)
    run_o(p)
    run_g(q)
    G.compare(q, p, TOL, fields)


def test_two_contexts_concurrently(G):
    """Two host threads, each with its own context (own stream) on the same GPU, run different problems at the same time
    (ctypes drops the GIL inside the library): no shared mutable state between contexts, results equal the oracle's."""
    import threading
    jobs = []
    for k, (L, M, H, kind) in enumerate([(90, 1500, 12, "sparse"), (130, 900, 20, "dense")]):
        Y = synth(L, M, H // 2, seed=900 + k)
        if kind == "sparse":
            p = vo.vbmf_sparse_init(Y, H, rng=np.random.default_rng(k))
        else:
            p = vo.vbmf_init(Y, H, rng=np.random.default_rng(k))
        jobs.append((kind, Y, p))
    outs, errs = [None, None], []

    def work(k):
        try:
            kind, Y, p = jobs[k]
            c = G.vb.Context(device=0)
            Yf = np.asfortranarray(Y)
            for rep in range(6):                      # several calls so the two threads really overlap
                q = G.to_gpu_params(p)
                if kind == "sparse":
                    G.vb.vbmf_sparse_(Yf, q, 8, eps=0.0, full_cov=True, ctx=c, yhat=False)
                else:
                    G.vb.vbmf_(Yf, q, 8, eps=0.0, est_covs=True, est_var=True, ctx=c, yhat=False)
            outs[k] = q
            c.close()
        except Exception as e:                        # surfaced in the main thread
            errs.append(e)
    ts = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for k, (kind, Y, p) in enumerate(jobs):
        po = copy.deepcopy(p)
        if kind == "sparse":
            vo.vbmf_sparse_run(Y, po, 8, eps=0.0, full_cov=True)
            floor = G.sensitivity(lambda s_: vo.vbmf_sparse_run(Y, s_, 8, eps=0.0, full_cov=True), p, ["AHat", "BHat"])
            G.compare(outs[k], po, max(TOL, 100 * floor), ["AHat", "BHat", "SigmaA", "SigmaB", "CA", "sigmaHat"])
        else:
            vo.vbmf_run(Y, po, 8, eps=0.0, est_covs=True, est_var=True)
            floor = G.sensitivity(lambda s_: vo.vbmf_run(Y, s_, 8, eps=0.0, est_covs=True, est_var=True), p, ["AHat", "BHat"])
            G.compare(outs[k], po, max(TOL, 100 * floor), ["AHat", "BHat", "SigmaA", "SigmaB", "CA", "CB", "sigma2"])


def test_checkpoint_resume(G, ctx, tmp_path):
    """SURVEY section 5, checkpoint / resume: the parameter struct IS the state.  (a) 4 + 6 iterations in two calls give what 10
    iterations in one call give (every field round-trips through the ABI; not bit for bit: on re-entry B'B is recomputed by the
    stand-alone Gram kernel, whose fixed summation order differs from the B epilogue's); (b) a state rebuilt from iteration 37 of a
    saved log (load_log + extract_params!, src/data_manip.jl:74-118) and continued for 63 iterations lands on the log's last entry."""
    Y = synth(60, 700, 4, seed=3)
    Yf = np.asfortranarray(Y)
    p = vo.vbmf_dual_init(Y, 8, 5, rng=np.random.default_rng(4))
    one, two = G.to_gpu_params(p), G.to_gpu_params(p)
    G.vb.vbmf_dual_(Yf, one, 10, eps=0.0, full_cov=True, ctx=ctx, yhat=False)
    G.vb.vbmf_dual_(Yf, two, 4, eps=0.0, full_cov=True, ctx=ctx, yhat=False)
    G.vb.vbmf_dual_(Yf, two, 6, eps=0.0, full_cov=True, ctx=ctx, yhat=False)
    for f in ("AHat", "BHat", "SigmaA", "SigmaB", "CA", "beta", "CB", "delta", "diagSigmaATVec"):
        assert G.rel(getattr(two, f), getattr(one, f)) < TOL, f
    for f in ("sigmaHat", "zeta", "alpha00", "alpha01", "beta00", "beta01"):
        assert abs(getattr(one, f) - getattr(two, f)) <= TOL * abs(getattr(one, f)), f
    g = load_golden("vbmf_test")
    Yg, pg = dense_state_from_golden(g)
    q = G.to_gpu_params(pg)
    G.vb.vbmf_(np.asfortranarray(Yg), q, 100, eps=0.0, est_covs=True, est_var=True, ctx=ctx, logdir=str(tmp_path), desc="ck")
    log, _ = G.vb.load_log(str(tmp_path / "ck"))
    r = G.to_gpu_params(pg)
    G.vb.extract_params_(log, r, 37)
    G.vb.vbmf_(np.asfortranarray(Yg), r, 63, eps=0.0, est_covs=True, est_var=True, ctx=ctx, yhat=False)
    for f in ("AHat", "BHat", "SigmaA", "SigmaB", "CA", "CB"):
        assert G.rel(getattr(r, f), log[f][..., 100]) < TOL, f
    assert abs(r.sigma2 - log["sigma2"][100]) <= TOL * abs(log["sigma2"][100])


# ------------------------------------------------------------------------------------------------ batched vbls!: diag_var and H > 32
@pytest.mark.parametrize("kind,full_cov,H,Mmax", [("sparse", True, 12, 30), ("sparse", False, 12, 30), ("dual", True, 20, 40), ("dual", False, 9, 25),
                                                  ("trial", True, 10, 20)])
def test_vbls_batched_diag_var(G, ctx, kind, full_cov, H, Mmax):
    """vbls!(Y, params, niter; diag_var = true, full_cov) (examples/mil_util.jl:179-203): heteroscedastic noise in the
    one-CTA-per-problem kernel -- B'diag(sv)B, B'diag(sv)Y and the per-row zeta / sigmaVecHat are per-iteration quantities."""
    rng = np.random.default_rng(23)
    L, nprob, niter = 17, 24, 8
    Ys, ps, qs = [], [], []
    for b in range(nprob):
        M = int(rng.integers(2, Mmax + 1))
        Y = 3.0 * synth(L, M, 3, seed=300 + b)
        if kind == "sparse":
            p = vo.vbmf_sparse_init(Y, H, rng=rng)
        elif kind == "trial":
            p = vo.vbmf_trial_init(Y, H, H - 2, M // 2, rng=rng)
        else:
            p = vo.vbmf_dual_init(Y, H, H - 1, rng=rng)
        p.SigmaB = np.diag(rng.uniform(1e-3, 1e-2, H))
        p.sigmaVecHat = rng.uniform(0.5, 2.0, L)
        Ys.append(np.asfortranarray(Y)); ps.append(p); qs.append(G.to_gpu_params(p))
    G.vb.vbls_batched_(Ys, qs, niter, full_cov=full_cov, diag_var=True, ctx=ctx)
    for Y, p, q in zip(Ys, ps, qs):
        vo.vbls(Y, p, niter, full_cov=full_cov, diag_var=True)
        G.compare(q, p, TOL, ["AHat", "ATVecHat", "diagSigmaATVec", "SigmaA", "CA", "beta", "sigmaVecHat", "zetaVec"])
        assert G.rel(q.YHat, p.YHat) < TOL


@pytest.mark.parametrize("kind,full_cov,diag_var,H", [("sparse", False, False, 48), ("sparse", True, False, 40), ("dual", True, True, 36), ("dual", False, False, 64)])
def test_vbls_batched_wide_rank(G, ctx, kind, full_cov, diag_var, H):
    """32 < H <= 64 in the batched path (full_cov: one column at a time, the whole CTA on one shared-memory matrix)."""
    rng = np.random.default_rng(29)
    L, nprob, niter = 20, 6, 5
    Ys, ps, qs = [], [], []
    for b in range(nprob):
        M = int(rng.integers(2, 13))
        Y = 2.0 * synth(L, M, 3, seed=700 + b)
        p = vo.vbmf_sparse_init(Y, H, rng=rng) if kind == "sparse" else vo.vbmf_dual_init(Y, H, H // 2, rng=rng)
        p.SigmaB = np.diag(rng.uniform(1e-3, 1e-2, H))
        Ys.append(np.asfortranarray(Y)); ps.append(p); qs.append(G.to_gpu_params(p))
    G.vb.vbls_batched_(Ys, qs, niter, full_cov=full_cov, diag_var=diag_var, ctx=ctx)
    for Y, p, q in zip(Ys, ps, qs):
        vo.vbls(Y, p, niter, full_cov=full_cov, diag_var=diag_var)
        G.compare(q, p, TOL, ["AHat", "diagSigmaATVec", "SigmaA", "CA", "beta", "sigmaHat", "sigmaVecHat"])
    with pytest.raises(G.vb.VBMFError, match="H <= 64|too large"):
        Yb = np.asfortranarray(synth(L, 5, 2, seed=1))
        G.vb.vbls_batched_([Yb], [G.to_gpu_params(vo.vbmf_sparse_init(Yb, 65, rng=rng))], 2, ctx=ctx)


# ------------------------------------------------------------------------------------------------ asynchronous chunked upload of Y
def test_chunked_upload_overlapped_first_iteration(G, monkeypatch):
    """attach_Y of a large matrix returns while the column chunks are still travelling; the first dense iteration follows the
    chunks (K1 / A epilogue / K2 per chunk), everything else waits for the upload.  Forced at a small size through
    VBMF_B200_ATTACH_CHUNK_MB; results must equal the oracle's (and the unchunked path's) for every kind."""
    L, M, H = 300, 4100, 16
    Y = synth(L, M, 8, seed=41)
    Yf = np.asfortranarray(Y)
    p = vo.vbmf_init(Y, H, H1=2, labels=list(range(5, M, 11)), rng=np.random.default_rng(3))
    ref = copy.deepcopy(p)
    vo.vbmf_run(Y, ref, 1, eps=0.0, est_covs=True, est_var=True)
    ref5 = copy.deepcopy(p)
    vo.vbmf_run(Y, ref5, 5, eps=0.0, est_covs=True, est_var=True)
    plain = G.to_gpu_params(p)
    c0 = G.vb.Context(device=0)
    G.vb.vbmf_(Yf, plain, 5, eps=0.0, est_covs=True, est_var=True, ctx=c0)
    c0.close()
    monkeypatch.setenv("VBMF_B200_ATTACH_CHUNK_MB", "1")          # 384-column chunks -> 11 chunks
    c = G.vb.Context(device=0)
    try:
        q1 = G.to_gpu_params(p)
        G.vb.vbmf_(Yf, q1, 1, eps=0.0, est_covs=True, est_var=True, ctx=c)       # the chunk-following iteration alone
        G.compare(q1, ref, TOL)
        q5 = G.to_gpu_params(p)
        G.vb.vbmf_(Yf, q5, 5, eps=0.0, est_covs=True, est_var=True, ctx=c)       # followed by ordinary iterations
        floor = G.sensitivity(lambda s_: vo.vbmf_run(Y, s_, 5, eps=0.0, est_covs=True, est_var=True), p, G.FIELDS["dense"])
        G.compare(q5, ref5, max(TOL, 100 * floor))
        G.compare(q5, plain, max(TOL, 100 * floor), G.FIELDS["dense"])
        assert G.rel(q5.YHat, q5.BHat @ q5.AHat.T) < 1e-13
        # the other consumers of Y wait for the upload: trYTY, the contractions, a sparse run
        c.attach(Yf)
        assert abs(c.trYTY() - float(np.sum(Y * Y))) <= 1e-12 * float(np.sum(Y * Y))
        c.attach(Yf)
        B = np.random.default_rng(0).standard_normal((L, H))
        assert G.rel(c.gemm_YtB(B), Y.T @ B) < 1e-13
        ps = vo.vbmf_sparse_init(Y, H, rng=np.random.default_rng(4))
        qs = G.to_gpu_params(ps)
        vo.vbmf_sparse_run(Y, ps, 2, eps=0.0, full_cov=False)
        G.vb.vbmf_sparse_(Yf, qs, 2, eps=0.0, full_cov=False, ctx=c)
        G.compare(qs, ps, TOL)
        # odd L (pitched copy) and a last chunk that is not a multiple of the tile
        Y2 = synth(301, 1000, 4, seed=42)
        p2 = vo.vbmf_init(Y2, 8, rng=np.random.default_rng(5))
        q2 = G.to_gpu_params(p2)
        vo.vbmf_run(Y2, p2, 2, eps=0.0, est_covs=True, est_var=True)
        G.vb.vbmf_(np.asfortranarray(Y2), q2, 2, eps=0.0, est_covs=True, est_var=True, ctx=c)
        G.compare(q2, p2, TOL)
    finally:
        c.close()
