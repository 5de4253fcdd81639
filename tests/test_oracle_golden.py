"""Pins the CPU oracle against the reference's own saved 100-iteration trajectories.

Free-running (not teacher-forced): start from logged slice 0, run 100 iterations, compare every logged field of
every slice.  These are the only result-pinning artefacts the reference ships (test/runtests.jl is a placeholder).
"""
import numpy as np
import pytest

from oracle import vbmf_oracle as vo
from tests.helpers import dense_state_from_golden, jl, load_golden, relerr, sparse_state_from_golden

TOL = 5e-12


def test_golden_checkpoints():
    g = load_golden("vbmf_test")
    Y = g["Y"].T
    assert Y.shape == (10, 20)
    assert Y[0, 0] == 1.2250170051991993 and Y[9, 19] == -0.30562420540482405
    assert abs((Y ** 2).sum() - 189.80687013285842) < 1e-12
    assert g["log_sigma2"][1] == 0.6259945822471235
    assert g["log_sigma2"][100] == 0.0023457154169626905
    assert (int(g["log_JULIA_MAJOR"][0]), int(g["log_JULIA_MINOR"][0])) == (0, 5)  # norm(Matrix) is spectral there
    s = load_golden("sparse_test")
    assert s["log_sigmaHat"][1] == 0.5324138163506104
    assert s["log_zeta"][100] == 21.72598805535064
    assert float(s["SigmaATVec_offblock_max"]) == 0.0


def test_dense_trajectory():
    g = load_golden("vbmf_test")
    Y, p = dense_state_from_golden(g)
    worst = {}

    def trace(p, i):
        for f in ("AHat", "BHat", "SigmaA", "SigmaB", "CA", "CB", "invCA", "invCB"):
            worst[f] = max(worst.get(f, 0.0), relerr(getattr(p, f), jl(g["log_" + f], i)))
        worst["sigma2"] = max(worst.get("sigma2", 0.0), abs(p.sigma2 - g["log_sigma2"][i]) / g["log_sigma2"][i])

    _, iters, d = vo.vbmf_run(Y, p, 100, eps=1e-6, est_covs=True, est_var=True, trace=trace)
    assert iters == 100 and d > 1e-6
    assert max(worst.values()) < TOL, worst
    # YHat is logged before its post-loop refresh: every slice holds the initial BHat*AHat'
    assert relerr(jl(g["log_BHat"], 0) @ jl(g["log_AHat"], 0).T, jl(g["log_YHat"], 0)) < 1e-14


def test_dense_sigma2_literal_identity():
    g = load_golden("vbmf_test")
    Y, p = dense_state_from_golden(g, 50)
    import copy
    q = copy.deepcopy(p)
    vo.dense_updateSigma2(Y, p, literal=True)
    vo.dense_updateSigma2(Y, q, literal=False)
    assert abs(p.sigma2 - q.sigma2) <= 1e-13 * abs(p.sigma2)
    assert abs(p.sigma2 - g["log_sigma2"][50]) <= 1e-10 * p.sigma2


@pytest.mark.parametrize("literal", [False, True])
def test_sparse_trajectory(literal):
    g = load_golden("sparse_test")
    Y, p = sparse_state_from_golden(g)
    worst = {}
    niter = 100 if not literal else 10

    def trace(p, i):
        for f in ("AHat", "BHat", "SigmaA", "SigmaB"):
            worst[f] = max(worst.get(f, 0.0), relerr(getattr(p, f), jl(g["log_" + f], i)))
        for f in ("ATVecHat", "diagSigmaATVec", "CA", "beta", "CB", "delta"):
            worst[f] = max(worst.get(f, 0.0), relerr(getattr(p, f), g["log_" + f][i]))
        for f in ("sigmaHat", "zeta"):
            worst[f] = max(worst.get(f, 0.0), abs(getattr(p, f) - g["log_" + f][i]) / abs(g["log_" + f][i]))
        worst["blocks"] = max(worst.get("blocks", 0.0), relerr(p.SigmaATVec_blocks, g["log_SigmaATVec"][i]))

    # the literal (MH)x(MH) formula and the block-wise evaluation must both track the log
    if literal:
        old = p.BHat.copy()
        for i in range(1, niter + 1):
            vo.sparse_updateA(Y, p, full_cov=True, literal=True)
            vo.sparse_updateB(Y, p)
            vo.sparse_updateCA(p)
            vo.sparse_updateCB(p)
            vo.sparse_updateSigma(Y, p)
            trace(p, i)
    else:
        d, iters = vo.vbmf_sparse_run(Y, p, niter, eps=1e-6, full_cov=True, diag_var=False, est_cb=True, trace=trace)
        assert iters == 100
        assert abs(p.CB[0] - 4.500000000493e10) / 4.5e10 < 1e-9
    assert max(worst.values()) < TOL, worst


def test_q2_repeat_inner():
    """Quirk Q2: the tail of the diagonal precision is repeat(d, inner=M-1), not a tiling."""
    Y = np.arange(12.0).reshape(3, 4)
    p = vo.vbmf_sparse_init(Y, 2, rng=np.random.default_rng(1))
    p.SigmaB = np.diag([0.5, 0.25])
    d = vo._diag_precision_base(p, False)
    vo.sparse_updateA(Y, p, full_cov=False)
    prec = 1.0 / p.diagSigmaATVec - 1.0  # CA = 1
    M, H = 4, 2
    expect = np.concatenate([d, np.repeat(d, M - 1)])
    assert np.allclose(prec, expect, rtol=1e-13)
    # closed form used by the CUDA path: j>H (1-based) -> d[ceil((j-H)/(M-1))]
    j = np.arange(1, M * H + 1)
    src = np.where(j <= H, j, np.ceil((j - H) / (M - 1))).astype(int) - 1
    assert np.array_equal(expect, d[src])


def test_preprocess_guards():
    """src/util.jl:36-87: constant rows survive scaleY as zeros and are dropped by preprocess; lambda multiplies the rest."""
    rng = np.random.default_rng(0)
    Y = rng.standard_normal((6, 50))
    Y[2, :] = 4.0
    out, used = vo.preprocess(Y, 10.0)
    assert list(used) == [1, 2, 4, 5, 6] and out.shape == (5, 50)
    s = out / 10.0
    assert np.allclose(s.mean(axis=1), 0.0, atol=1e-12) and np.allclose(s.var(axis=1, ddof=1), 1.0)
