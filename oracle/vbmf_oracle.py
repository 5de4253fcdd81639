"""NumPy restatement of the reference's VB matrix-factorisation update loop (CPU, Float64).

TEST INFRASTRUCTURE ONLY -- the checker, never the thing measured or shipped.  The product path
(vbmatrixfactorization.jl_b200/) never imports this module and has no CPU fallback.

Parity status
-------------
PINNED against the reference's own artefacts (tests/test_oracle_golden.py):
  * dense `vbmf` (est_covs, est_var)            <- /root/reference/examples/data/vbmf_test/log.jld   (100 iterations)
  * `vbmf_sparse` full_cov, homoscedastic, est_cb <- /root/reference/examples/data/sparse_test/log.jld (100 iterations)
  Both are replayed free-running from the logged initial state and must match every logged field.
PARITY UNPINNED (no reference fixture or test exercises them; this restatement is the specification):
  full_cov=false (diagonal) path, diag_var=true, labels/H1 masking, all of vbmf_dual (incl. the
  Roots.jl `fzero` root, a third-party dependency that is neither vendored nor listed in REQUIRE),
  lowerBound / lowerBoundTrimmed, early convergence exit.
  These branches are cross-checked (tests/test_oracle_mp.py) against a second, independent restatement of the literal
  reference formulae in 40-digit arithmetic (oracle/mp_restatement.py): agreement <= 1e-12 per update step.  That guards
  against a transcription slip in this file; it does not replace a run of the reference itself.

Every function cites the reference file:line (relative to /root/reference) it follows.  Julia 0.5.2
semantics (the fixtures record JULIA 0.5.2) are honoured: column-major vec/reshape, `inv` = LU with partial
pivoting (LAPACK getrf/getri -> scipy.linalg.inv), `norm(::Matrix)` = spectral norm, `repeat(inner=)`.
Arrays keep the Julia shapes: AHat (M,H), BHat (L,H), ATVecHat (M*H,) with H fastest, labels 1-based.
Objects the reference materialises at O((MH)^2), O(M^2) or O((LH)^2) are evaluated block-wise / by trace
identities so the oracle also runs at bench sizes; `literal=True` switches select the literal formula for
tiny-size identity tests.
"""
from __future__ import annotations

import copy as _copy
import math
from types import SimpleNamespace

import numpy as np
import scipy.linalg as sla
from scipy.special import digamma, gammaln

ln2pi = math.log(2 * math.pi)  # src/util.jl:1
EPS0 = 4.9406564584124654e-324  # Julia eps(0.0), src/util.jl:120-122


# ----------------------------------------------------------------------------- util.jl
def norm2(x):
    """src/util.jl:8-19  sum of squares."""
    if hasattr(x, "norm2"):          # ContractedY (below): sum(Y.^2) supplied by the caller
        return float(x.norm2())
    return float(np.sum(np.asarray(x) ** 2))


class ContractedY:
    """Stand-in for Y at sizes where the CPU cannot form Y'*B and Y*A in reasonable time: the two contractions (and
    sum(Y.^2)) are supplied as callables, every other line of the update loop still runs in this oracle.  Used by the
    full-size GPU tests, where the contractions themselves are validated through size-independent properties."""

    __array_ufunc__ = None            # ndarray @ ContractedY defers to __rmatmul__ below

    def __init__(self, shape, ytb, ya, trYTY):
        self.shape = tuple(shape)
        self._ytb, self._ya, self._tr = ytb, ya, float(trYTY)

    def __rmatmul__(self, Bt):
        """B' * Y (H x M), the form src/vbmf_sparse.jl:195,232 uses: (Y' * B)'."""
        return np.asarray(self._ytb(np.ascontiguousarray(np.asarray(Bt).T))).T

    class _T:
        def __init__(self, outer):
            self.o = outer

        def __matmul__(self, B):
            return np.asarray(self.o._ytb(np.asarray(B)))

    @property
    def T(self):
        return ContractedY._T(self)

    def __matmul__(self, A):
        return np.asarray(self._ya(np.asarray(A)))

    def norm2(self):
        return self._tr


def matnorm(X, mode="spectral"):
    """Julia 0.5 `norm(::Matrix)` = largest singular value (Quirk Q1); `frobenius` = Julia >= 0.7 reading."""
    X = np.asarray(X)
    if X.ndim == 1:
        return float(np.linalg.norm(X))
    if mode == "spectral":
        return float(np.linalg.norm(X, 2))
    if mode == "frobenius":
        return float(np.linalg.norm(X, "fro"))
    raise ValueError(mode)


def delta(new, old, mode="spectral"):
    """src/util.jl:27-29  norm(old - new)/norm(old)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.float64(matnorm(old - new, mode)) / np.float64(matnorm(old, mode)))


def scaleY(Y):
    """src/util.jl:36-53: row-standardise (corrected variance), numerical-zero guards 1e-15 / 1e-8."""
    Y = np.asarray(Y, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        mu = Y.mean(axis=1, keepdims=True)
        den = Y.var(axis=1, ddof=1, keepdims=True)
        den[np.abs(den) <= 1e-15] = 1.0
        den[den == 0.0] = 1.0
        den = np.sqrt(den)
        nom = Y - mu
        nom[np.abs(nom) <= 1e-8] = 0.0
        return nom / den


def preprocess(Y, lam):
    """src/util.jl:73-87: scaleY, drop rows with sum(abs) < 1e-5, multiply by lambda.  Returns (Y_out, used_rows 1-based)."""
    sY = scaleY(Y)
    rowsums = np.sum(np.abs(sY), axis=1)
    used = np.nonzero(rowsums >= 1e-5)[0]
    return lam * sY[used, :], used + 1


def traceXTY(X, Y):
    """src/util.jl:104-106."""
    if X is Y and hasattr(X, "norm2"):      # traceXTY(Y, Y) of a ContractedY
        return float(X.norm2())
    return float(np.sum(X * Y))


def normalEntropy_diag(diagSigma):
    """src/util.jl:131-137."""
    n = diagSigma.shape[0]
    return n / 2 + n / 2 * ln2pi + 0.5 * float(np.sum(np.log(diagSigma)))


def normalEntropy_mat(Sigma):
    """src/util.jl:113-125 literal: det by LU, clamped to eps(0.0)."""
    m = Sigma.shape[0]
    d = float(np.linalg.det(Sigma))
    if d < EPS0:
        d = EPS0
    return m / 2 + m / 2 * ln2pi + 0.5 * math.log(d)


def normalEntropy_kronSigmaB(SigmaB, L):
    """normalEntropy(kron(SigmaB, eye(L))) (src/vbmf_sparse.jl:463, src/util.jl:113-125) without the (LH)^2 matrix.

    LU with partial pivoting of kron(SigmaB, I_L) has U-diagonal u_11 (L times), u_22 (L times), ... where u_aa is the
    U-diagonal of lu(SigmaB); Julia's det multiplies them left to right in Float64, so the running product
    saturates at 0 / Inf (Quirk Q7).  Emulated in the log domain at the ends of each run of L equal factors.
    """
    H = SigmaB.shape[0]
    m = L * H
    lu, piv = sla.lu_factor(SigmaB)
    u = np.diag(lu)
    sign = 1.0
    for i, p in enumerate(piv):
        if p != i:
            sign = -sign
    sign = sign ** L if L % 2 else 1.0
    LOG_MIN = math.log(EPS0) - math.log(2.0)  # below this the running product rounds to 0
    LOG_MAX = math.log(np.finfo(np.float64).max)
    run = 0.0
    state = "finite"
    for a in range(H):
        if u[a] == 0.0:
            state = "zero"
            break
        if u[a] < 0:
            sign = -sign if L % 2 else sign
        run += L * math.log(abs(u[a]))
        if run < LOG_MIN:
            state = "zero"
            break
        if run > LOG_MAX:
            state = "inf"
            break
    if state == "zero":
        logd = math.log(EPS0)
    elif state == "inf":
        logd = math.inf if sign > 0 else math.log(EPS0)
    else:
        logd = run if sign > 0 else math.log(EPS0)
        if logd < math.log(EPS0):
            logd = math.log(EPS0)
    return m / 2 + m / 2 * ln2pi + 0.5 * logd


def gammaEntropy(a, b):
    """src/util.jl:144-146  (a + log(b) + lgamma(a) + (1-a)*digamma(a), '+log b' as written)."""
    return a + np.log(b) + gammaln(a) + (1 - a) * digamma(a)


def gammaELn(a, b):
    """src/util.jl:153-155."""
    return digamma(a) - np.log(b)


def _inv(A):
    """Julia `inv` -> LAPACK getrf + getri."""
    return sla.inv(A, check_finite=False)


def _labels0(p):
    """1-based Julia labels -> 0-based numpy row indices."""
    return np.asarray(p.labels, dtype=np.int64) - 1


def _mask(p):
    """AHat[labels, end-H1+1:end] = 0.0  (src/vbmf.jl:101, src/vbmf_sparse.jl:245)."""
    if p.H1 > 0 and len(p.labels) > 0:
        p.AHat[_labels0(p), p.H - p.H1:] = 0.0


# ----------------------------------------------------------------------------- dense vbmf (src/vbmf.jl)
def vbmf_init(Y, H, ca=1.0, cb=1.0, sigma2=1.0, H1=0, labels=(), rng=None, AHat=None, BHat=None):
    """src/vbmf.jl:48-73.  The RNG stream is Julia's in the reference; here AHat/BHat may be supplied."""
    rng = rng or np.random.default_rng(0)
    L, M = Y.shape
    p = SimpleNamespace(kind="dense")
    p.L, p.M, p.H, p.H1 = L, M, H, H1
    p.labels = np.asarray(labels, dtype=np.int64)
    p.AHat = rng.standard_normal((M, H)) if AHat is None else np.array(AHat, dtype=np.float64)  # A before B (:58,:62)
    _mask(p)
    p.BHat = rng.standard_normal((L, H)) if BHat is None else np.array(BHat, dtype=np.float64)
    p.SigmaA = np.zeros((H, H))
    p.SigmaB = np.zeros((H, H))
    p.CA = ca * np.eye(H)
    p.CB = cb * np.eye(H)
    p.invCA = _inv(p.CA)
    p.invCB = _inv(p.CB)
    p.sigma2 = float(sigma2)
    p.YHat = None  # lazily: BHat @ AHat.T (:70)
    return p


def dense_updateA(Y, p):
    """src/vbmf.jl:95-102."""
    p.SigmaA = p.sigma2 * _inv(p.BHat.T @ p.BHat + p.L * p.SigmaB + p.sigma2 * p.invCA)
    p.AHat = ((Y.T @ p.BHat) @ p.SigmaA) / p.sigma2
    _mask(p)


def dense_updateB(Y, p):
    """src/vbmf.jl:109-113."""
    p.SigmaB = p.sigma2 * _inv(p.AHat.T @ p.AHat + p.M * p.SigmaA + p.sigma2 * p.invCB)
    p.BHat = ((Y @ p.AHat) @ p.SigmaB) / p.sigma2


def dense_updateCA(p):
    """src/vbmf.jl:129-134 (mutates CA in place -- Quirk Q5)."""
    for h in range(p.H):
        p.CA[h, h] = norm2(p.AHat[:, h]) / p.M + p.SigmaA[h, h]
    p.invCA = _inv(p.CA)


def dense_updateCB(p):
    """src/vbmf.jl:141-146."""
    for h in range(p.H):
        p.CB[h, h] = norm2(p.BHat[:, h]) / p.L + p.SigmaB[h, h]
    p.invCB = _inv(p.CB)


def dense_updateSigma2(Y, p, literal=False):
    """src/vbmf.jl:153-157.  literal=True forms the M x M product as the reference does."""
    GA = p.AHat.T @ p.AHat + p.M * p.SigmaA
    GB = p.BHat.T @ p.BHat + p.L * p.SigmaB
    if literal:
        cross = float(np.trace(2 * Y.T @ p.BHat @ p.AHat.T))
    else:
        cross = 2.0 * traceXTY(p.BHat, Y @ p.AHat)  # tr(Y' B A') = sum(B .* (Y A))
    p.sigma2 = (norm2(Y) - cross + float(np.trace(GA @ GB))) / (p.L * p.M)


def dense_updateYHat(p):
    """src/vbmf.jl:120-122."""
    p.YHat = p.BHat @ p.AHat.T


def vbmf_run(Y, p, niter, eps=1e-6, est_covs=False, est_var=False, norm="spectral", trace=None, yhat=True):
    """`vbmf!` src/vbmf.jl:175-231.  Returns (params, iterations_done, d).  trace(p, i) is called once per iteration."""
    old = p.BHat
    d = eps + 1.0
    i = 1
    while i <= niter and d > eps:
        dense_updateA(Y, p)
        dense_updateB(Y, p)
        if est_covs:
            dense_updateCA(p)
            dense_updateCB(p)
        if est_var:
            dense_updateSigma2(Y, p)
        if trace is not None:
            trace(p, i)
        d = delta(p.BHat, old, norm)
        old = p.BHat
        i += 1
    if yhat:        # the L x M product is skipped by the full-size tests
        dense_updateYHat(p)
    return p, i - 1, d


def vbmf(Y, p_in, niter, **kw):
    """src/vbmf.jl:238-248.  The reference's copy is shallow (Q5); results are identical, so deep-copy here."""
    return vbmf_run(Y, _copy.deepcopy(p_in), niter, **kw)


# ----------------------------------------------------------------------------- vbmf_sparse (src/vbmf_sparse.jl)
def vbmf_sparse_init(Y, H, ca=1.0, alpha0=1e-10, beta0=1e-10, cb=1.0, gamma0=1e-10, delta0=1e-10, sigma=1.0,
                     eta0=1e-10, zeta0=1e-10, H1=0, labels=(), rng=None, AHat=None, BHat=None):
    """src/vbmf_sparse.jl:101-153.  SigmaATVec/invSigmaATVec (eye(MH,MH), :120,:122) are kept as M identity blocks."""
    rng = rng or np.random.default_rng(0)
    L, M = Y.shape
    p = SimpleNamespace(kind="sparse")
    p.L, p.M, p.H, p.MH, p.H1 = L, M, H, M * H, H1
    p.labels = np.asarray(labels, dtype=np.int64)
    p.AHat = rng.standard_normal((M, H)) if AHat is None else np.array(AHat, dtype=np.float64)
    _mask(p)
    p.ATVecHat = p.AHat.reshape(M * H).copy()
    p.SigmaATVec_blocks = None  # M x H x H, only produced by the full_cov path
    p.diagSigmaATVec = np.ones(M * H)
    p.SigmaA = np.zeros((H, H))
    p.BHat = rng.standard_normal((L, H)) if BHat is None else np.array(BHat, dtype=np.float64)
    p.SigmaB = np.zeros((H, H))
    p.CA = ca * np.ones(M * H)
    p.alpha0, p.beta0 = alpha0, beta0
    p.alpha = alpha0 + 0.5
    p.beta = beta0 * np.ones(M * H)
    p.CB = cb * np.ones(H)
    p.gamma0, p.delta0 = gamma0, delta0
    p.gamma = gamma0 + L / 2
    p.delta = delta0 * np.ones(H)
    p.sigmaHat = float(sigma)
    p.eta0, p.zeta0 = eta0, zeta0
    p.eta = eta0 + L * M / 2
    p.zeta = zeta0
    p.sigmaVecHat = sigma * np.ones(L)
    p.etaVec = (eta0 + M / 2) * np.ones(L)
    p.zetaVec = zeta0 * np.ones(L)
    p.YHat = None
    p.trYTY = traceXTY(Y, Y)
    return p


def _diag_precision_base(p, diag_var):
    """First H entries of the diagonal precision, src/vbmf_sparse.jl:207-219 (Q3, Q4)."""
    d = np.empty(p.H)
    if diag_var:
        ms = float(np.mean(p.sigmaVecHat))
        for h in range(p.H):
            d[h] = norm2(p.BHat[:, h] * p.sigmaVecHat) + p.L * ms * p.SigmaB[h, h]
    else:
        for h in range(p.H):
            d[h] = p.sigmaHat * norm2(p.BHat[:, h]) + p.L * p.SigmaB[h, h]
    return d


def _sparse_updateA_core(Y, p, full_cov, diag_var, literal=False):
    """Body of updateA! up to (not including) the AHat reshape/mask: src/vbmf_sparse.jl:176-240 == src/vbmf_dual.jl:216-280."""
    L, M, H = p.L, p.M, p.H
    if full_cov:
        if diag_var:
            G = p.BHat.T @ (p.sigmaVecHat[:, None] * p.BHat) + L * float(np.mean(p.sigmaVecHat)) * p.SigmaB
            V = p.BHat.T @ (p.sigmaVecHat[:, None] * Y)          # H x M ; vec(V) = reshape(B' diag(s) Y, HM)
        else:
            G = p.sigmaHat * (p.BHat.T @ p.BHat + L * p.SigmaB)
            V = p.BHat.T @ Y
        if literal:  # the reference's (MH)x(MH) formula, :181-195
            invS = np.kron(np.eye(M), G) + np.diag(p.CA)
            S = _inv(invS)
            p.diagSigmaATVec = np.diag(S).copy()
            v = V.T.reshape(M * H)
            p.ATVecHat = S @ v if diag_var else (p.sigmaHat * S) @ v
            blocks = np.stack([S[m * H:(m + 1) * H, m * H:(m + 1) * H] for m in range(M)])
        else:
            blocks = np.empty((M, H, H))
            avec = np.empty((M, H))
            CA = p.CA.reshape(M, H)
            for m in range(M):
                Sm = _inv(G + np.diag(CA[m]))
                blocks[m] = Sm
                avec[m] = Sm @ V[:, m] if diag_var else (p.sigmaHat * Sm) @ V[:, m]
            p.diagSigmaATVec = np.einsum("mhh->mh", blocks).reshape(M * H).copy()
            p.ATVecHat = avec.reshape(M * H)
        p.SigmaATVec_blocks = blocks
        SA = np.zeros((H, H))
        for m in range(M):  # :198-201 sequential accumulation
            SA += blocks[m]
        p.SigmaA = SA
    else:
        d = _diag_precision_base(p, diag_var)
        prec = np.empty(M * H)
        prec[:H] = d
        prec[H:] = np.repeat(d, M - 1)          # Quirk Q2: repeat(...; inner = M-1), :221
        prec = prec + p.CA                       # :223
        s = 1.0 / prec                           # :226
        p.diagSigmaATVec = s
        if diag_var:
            v = (p.BHat.T @ (p.sigmaVecHat[:, None] * Y)).T.reshape(M * H)
            p.ATVecHat = s * v                   # :230
        else:
            v = (p.BHat.T @ Y).T.reshape(M * H)
            p.ATVecHat = (p.sigmaHat * s) * v    # :232  (left-to-right)
        p.SigmaA = np.diag(s.reshape(M, H).sum(axis=0))   # :236-239
        p.SigmaATVec_blocks = None


def sparse_updateA(Y, p, full_cov=False, diag_var=False, literal=False):
    """src/vbmf_sparse.jl:176-247."""
    _sparse_updateA_core(Y, p, full_cov, diag_var, literal)
    p.AHat = p.ATVecHat.reshape(p.M, p.H).copy()      # reshape(ATVecHat, H, M)'
    _mask(p)
    p.ATVecHat = p.AHat.reshape(p.M * p.H).copy()


def sparse_updateB(Y, p, diag_var=False):
    """src/vbmf_sparse.jl:254-268 == src/vbmf_dual.jl:292-306."""
    GA = p.AHat.T @ p.AHat + p.SigmaA
    if diag_var:
        p.SigmaB = _inv(np.diag(p.CB) + float(np.mean(p.sigmaVecHat)) * GA)
        p.BHat = (p.sigmaVecHat[:, None] * (Y @ p.AHat)) @ p.SigmaB      # diagm(sv)*Y*AHat*SigmaB
    else:
        p.SigmaB = _inv(np.diag(p.CB) + p.sigmaHat * GA)
        p.BHat = (p.sigmaHat * (Y @ p.AHat)) @ p.SigmaB                  # ((sigmaHat*Y)*AHat)*SigmaB, Q12


def sparse_updateCA(p):
    """src/vbmf_sparse.jl:284-288."""
    p.beta = p.beta0 * np.ones(p.M * p.H) + 0.5 * (p.ATVecHat * p.ATVecHat + p.diagSigmaATVec)
    p.CA = p.alpha * np.ones(p.M * p.H) / p.beta


def sparse_updateCB(p):
    """src/vbmf_sparse.jl:295-300 (Q6: 1/2*SigmaB[h,h], not L/2)."""
    for h in range(p.H):
        p.delta[h] = p.delta0 + 0.5 * float(p.BHat[:, h] @ p.BHat[:, h]) + 0.5 * p.SigmaB[h, h]
        p.CB[h] = p.gamma / p.delta[h]


def sparse_updateSigma(Y, p, diag_var=False):
    """src/vbmf_sparse.jl:307-323 == src/vbmf_dual.jl:370-386.  Y*AHat is formed once (the reference recomputes it)."""
    GA = p.AHat.T @ p.AHat + p.SigmaA
    Q = Y @ p.AHat
    if diag_var:
        for l in range(p.L):
            b = p.BHat[l, :]
            p.zetaVec[l] = (p.zeta0 + 0.5 * norm2(Y[l, :]) - float(np.sum(Q[l, :] * b))
                            + 0.5 * traceXTY(GA, np.outer(b, b) + p.SigmaB))
            p.sigmaVecHat[l] = p.etaVec[l] / p.zetaVec[l]
    else:
        p.zeta = (p.zeta0 + 0.5 * p.trYTY - traceXTY(p.BHat, Q)
                  + 0.5 * traceXTY(GA, p.BHat.T @ p.BHat + p.L * p.SigmaB))
        p.sigmaHat = p.eta / p.zeta


def sparse_updateYHat(p):
    """src/vbmf_sparse.jl:275-277."""
    p.YHat = p.BHat @ p.AHat.T


def vbmf_sparse_run(Y, p, niter, eps=1e-6, diag_var=False, full_cov=False, est_cb=True, norm="spectral", trace=None, yhat=True):
    """`vbmf_sparse!` src/vbmf_sparse.jl:344-410.  Returns (d, iterations_done)."""
    old = p.BHat.copy()
    d = eps + 1.0
    i = 1
    while i <= niter and d > eps:
        sparse_updateA(Y, p, full_cov=full_cov, diag_var=diag_var)
        sparse_updateB(Y, p, diag_var=diag_var)
        sparse_updateCA(p)
        if est_cb:
            sparse_updateCB(p)
        sparse_updateSigma(Y, p, diag_var=diag_var)
        if trace is not None:
            trace(p, i)
        d = delta(p.BHat, old, norm)
        old = p.BHat.copy()
        i += 1
    if yhat:        # the L x M product is skipped by the full-size tests
        sparse_updateYHat(p)
    return d, i - 1


def vbmf_sparse(Y, p_in, niter, **kw):
    """src/vbmf_sparse.jl:418-428 -> (params, d)."""
    p = _copy.deepcopy(p_in)
    d, _ = vbmf_sparse_run(Y, p, niter, **kw)
    return p, d


def _lb_common(Y, p):
    """Terms shared by the sparse and dual lower bounds: src/vbmf_sparse.jl:438-441 == src/vbmf_dual.jl:559-562."""
    GA = p.AHat.T @ p.AHat + p.SigmaA
    GB = p.BHat.T @ p.BHat + p.L * p.SigmaB
    Lb = -p.L * p.M / 2 * ln2pi + p.L * p.M / 2 * gammaELn(p.eta, p.zeta)
    Lb += -p.sigmaHat / 2 * (p.trYTY - 2 * traceXTY(p.BHat, Y @ p.AHat) + traceXTY(GA, GB))
    return Lb, GB


def sparse_lowerBound(Y, p):
    """src/vbmf_sparse.jl:435-471, term by term in the reference's order (Q7, Q8, Q13)."""
    Lb, GB = _lb_common(Y, p)
    eln_b = float(np.sum(gammaELn(p.alpha * np.ones(p.beta.shape), p.beta)))
    eln_d = float(np.sum(gammaELn(p.gamma * np.ones(p.delta.shape), p.delta)))
    # E[lnp(vec(A'))]
    Lb += -p.MH / 2 * ln2pi + 0.5 * eln_b
    Lb += -(0.5 * float(p.CA @ (p.ATVecHat ** 2 + p.diagSigmaATVec)))
    # E[lnp(B)]
    Lb += -p.L * p.H / 2 * ln2pi
    Lb += p.L / 2 * eln_d
    Lb += -0.5 * traceXTY(np.diag(p.CB), GB)
    # E[lnp(sigma)]
    Lb += p.eta0 * math.log(p.zeta0) - gammaln(p.eta0)
    Lb += (p.eta0 - 1) * gammaELn(p.eta, p.zeta) - p.zeta0 * p.sigmaHat
    # E[lnp(CA)]
    Lb += p.MH * (p.alpha0 * math.log(p.beta0) - gammaln(p.alpha0))
    Lb += (p.alpha0 - 1) * eln_b
    Lb += -p.beta0 * float(np.sum(p.CA))
    # E[lnp(CB)]
    Lb += p.H * (p.gamma0 * math.log(p.delta0) - gammaln(p.gamma0))
    Lb += (p.gamma0 - 1) * eln_d
    Lb += -p.gamma0 * float(np.sum(p.CB))       # Q13: gamma0, as written (:458)
    # entropies
    Lb += normalEntropy_diag(p.diagSigmaATVec)
    Lb += normalEntropy_kronSigmaB(p.SigmaB, p.L)
    Lb += gammaEntropy(p.eta, p.zeta)
    Lb += float(np.sum(gammaEntropy(p.alpha * np.ones(p.beta.shape), p.beta)))
    Lb += float(np.sum(gammaEntropy(p.gamma * np.ones(p.delta.shape), p.delta)))
    return float(Lb)


def _trim(p_in, trim):
    """src/vbmf_sparse.jl:478-487 == src/vbmf_dual.jl:606-615 (only these four vectors and MH change)."""
    p = _copy.deepcopy(p_in)
    keep = np.abs(p.ATVecHat) > trim
    p.ATVecHat = p.ATVecHat[keep]
    p.MH = int(p.ATVecHat.shape[0])
    p.beta = p.beta[keep]
    p.CA = p.CA[keep]
    p.diagSigmaATVec = p.diagSigmaATVec[keep]
    return p


def sparse_lowerBoundTrimmed(Y, p_in, trim=1e-1):
    """src/vbmf_sparse.jl:478-489."""
    return sparse_lowerBound(Y, _trim(p_in, trim))


# ----------------------------------------------------------------------------- vbmf_dual (src/vbmf_dual.jl)
def _interleave(p, v0, v1):
    """Per-row interleave [v0 block m ; v1 block m] built by the O(M^2) cat loops, src/vbmf_dual.jl:155-158,340-350."""
    M, H0, H1 = p.M, p.H0, p.H1
    out = np.empty((M, H0 + H1))
    out[:, :H0] = v0.reshape(M, H0)
    out[:, H0:] = v1.reshape(M, H1)
    return out.reshape(M * (H0 + H1))


def vbmf_dual_init(Y, H, H0, ca=1.0, alpha0=1e-10, beta0=1e-10, cb=1.0, gamma0=1e-10, delta0=1e-10, sigma=1.0,
                   eta0=1e-10, zeta0=1e-10, rng=None, AHat=None, BHat=None):
    """src/vbmf_dual.jl:122-193."""
    if H < H0:
        raise ValueError("H must be at least H0!")   # :126-128
    rng = rng or np.random.default_rng(0)
    L, M = Y.shape
    p = SimpleNamespace(kind="dual")
    p.L, p.M, p.H, p.MH, p.H0 = L, M, H, M * H, H0
    H1 = H - H0
    p.H1 = H1
    p.labels = np.zeros(0, dtype=np.int64)
    p.AHat = rng.standard_normal((M, H)) if AHat is None else np.array(AHat, dtype=np.float64)
    p.ATVecHat = p.AHat.reshape(M * H).copy()
    p.SigmaATVec_blocks = None
    p.diagSigmaATVec = np.ones(M * H)
    p.SigmaA = np.zeros((H, H))
    p.A0Hat = p.AHat[:, :H0].copy()
    p.A1Hat = p.AHat[:, H0:].copy()
    p.BHat = rng.standard_normal((L, H)) if BHat is None else np.array(BHat, dtype=np.float64)
    p.SigmaB = np.zeros((H, H))
    p.CA0 = ca * np.ones(M * H0)
    p.CA1 = ca * np.ones(M * H1)
    p.CA = _interleave(p, p.CA0, p.CA1)
    p.alpha00 = p.alpha01 = alpha0
    p.beta00 = p.beta01 = beta0
    p.alpha0 = alpha0 + 0.5
    p.beta0 = beta0 * np.ones(M * H0)
    p.alpha1 = alpha0 + 0.5
    p.beta1 = beta0 * np.ones(M * H1)
    p.alpha = np.array([p.alpha0, p.alpha1])
    p.beta = _interleave(p, p.beta0, p.beta1)
    p.CB = cb * np.ones(H)
    p.gamma0, p.delta0 = gamma0, delta0
    p.gamma = gamma0 + L / 2
    p.delta = delta0 * np.ones(H)
    p.sigmaHat = float(sigma)
    p.eta0, p.zeta0 = eta0, zeta0
    p.eta = eta0 + L * M / 2
    p.zeta = zeta0
    p.sigmaVecHat = sigma * np.ones(L)
    p.etaVec = (eta0 + M / 2) * np.ones(L)
    p.zetaVec = zeta0 * np.ones(L)
    p.YHat = None
    p.trYTY = traceXTY(Y, Y)
    return p


def dual_updateA(Y, p, full_cov=False, diag_var=False, literal=False):
    """src/vbmf_dual.jl:216-285 (no label mask; refreshes A0Hat/A1Hat)."""
    _sparse_updateA_core(Y, p, full_cov, diag_var, literal)
    p.AHat = p.ATVecHat.reshape(p.M, p.H).copy()
    p.A0Hat = p.AHat[:, :p.H0].copy()
    p.A1Hat = p.AHat[:, p.H0:].copy()


dual_updateB = sparse_updateB          # src/vbmf_dual.jl:292-306
dual_updateCB = sparse_updateCB        # src/vbmf_dual.jl:358-363
dual_updateSigma = sparse_updateSigma  # src/vbmf_dual.jl:370-386
dual_updateYHat = sparse_updateYHat    # src/vbmf_dual.jl:313-315


def dual_updateCA(p):
    """src/vbmf_dual.jl:322-351."""
    M, H, H0, H1 = p.M, p.H, p.H0, p.H1
    p.alpha0 = p.alpha00 + 0.5
    p.alpha1 = p.alpha01 + 0.5
    dS = p.diagSigmaATVec.reshape(M, H)          # reshape(diag, H, M) column m == row m here
    p.beta0 = p.beta00 * np.ones(M * H0) + 0.5 * (p.A0Hat * p.A0Hat + dS[:, :H0]).reshape(M * H0)
    p.beta1 = p.beta01 * np.ones(M * H1) + 0.5 * (p.A1Hat * p.A1Hat + dS[:, H0:]).reshape(M * H1)
    p.CA0 = p.alpha0 * np.ones(M * H0) / p.beta0
    p.CA1 = p.alpha1 * np.ones(M * H1) / p.beta1
    p.CA = _interleave(p, p.CA0, p.CA1)
    p.alpha = np.array([p.alpha0, p.alpha1])
    p.beta = _interleave(p, p.beta0, p.beta1)


def fzero_bisect(f, a=1e-10, b=1e10):
    """Stand-in for Roots.jl `fzero(f, 1e-10, 1e10, ftol=1e-5)` (src/vbmf_dual.jl:398,422).  PARITY UNPINNED.

    Roots' Float64 bracketing method bisects until the bracket cannot shrink, i.e. returns the root of the
    floating-point function to full precision irrespective of ftol; no sign change -> exception (Q11).
    """
    fa, fb = f(a), f(b)
    if not (np.isfinite(fa) and np.isfinite(fb)) or fa * fb > 0:
        raise ValueError("no bracket")
    if fa == 0:
        return a
    if fb == 0:
        return b
    for _ in range(4000):
        m = 0.5 * (a + b)
        if m <= a or m >= b:
            break
        fm = f(m)
        if fm == 0:
            return m
        if (fm > 0) == (fa > 0):
            a, fa = m, fm
        else:
            b, fb = m, fm
    return a if abs(fa) <= abs(fb) else b


def _update_alpha0x(N, beta0x, alpha_g, beta_g, old):
    """src/vbmf_dual.jl:393-401 / :417-425.  f(x) = N*log(beta0x) - N*digamma(x) + sum(gammaELn(alpha_g, beta_g))."""
    S = float(np.sum(gammaELn(alpha_g * np.ones(beta_g.shape), beta_g)))
    with np.errstate(divide="ignore", invalid="ignore"):
        c = N * np.log(np.float64(beta0x))

    def f(x):
        return c - N * digamma(x) + S
    try:
        return float(fzero_bisect(f))
    except Exception:  # `try ... end`, Q11: keep the old value
        return old


def dual_updateAlpha00(p):
    p.alpha00 = _update_alpha0x(p.M * p.H0, p.beta00, p.alpha0, p.beta0, p.alpha00)


def dual_updateAlpha01(p):
    p.alpha01 = _update_alpha0x(p.M * p.H1, p.beta01, p.alpha1, p.beta1, p.alpha01)


def dual_updateBeta00(p):
    """src/vbmf_dual.jl:408-410."""
    with np.errstate(divide="ignore", invalid="ignore"):
        p.beta00 = float(np.float64(p.M * p.H0 * p.alpha00) / np.sum(p.CA0))


def dual_updateBeta01(p):
    """src/vbmf_dual.jl:432-434."""
    with np.errstate(divide="ignore", invalid="ignore"):
        p.beta01 = float(np.float64(p.M * p.H1 * p.alpha01) / np.sum(p.CA1))


def vbmf_dual_run(Y, p, niter, eps=1e-6, diag_var=False, full_cov=False, est_priors=True, est_cb=True,
                  norm="spectral", trace=None, yhat=True):
    """`vbmf_dual!` src/vbmf_dual.jl:455-530.  Returns (d, iterations_done)."""
    old = p.BHat.copy()
    d = eps + 1.0
    i = 1
    while i <= niter and d > eps:
        dual_updateA(Y, p, full_cov=full_cov, diag_var=diag_var)
        dual_updateB(Y, p, diag_var=diag_var)
        dual_updateCA(p)
        if est_cb:
            dual_updateCB(p)
        dual_updateSigma(Y, p, diag_var=diag_var)
        if est_priors:      # :491-497, alphas first (they see the old beta0x)
            dual_updateAlpha00(p)
            dual_updateAlpha01(p)
            dual_updateBeta00(p)
            dual_updateBeta01(p)
        if trace is not None:
            trace(p, i)
        d = delta(p.BHat, old, norm)
        old = p.BHat.copy()
        i += 1
    if yhat:        # the L x M product is skipped by the full-size tests
        dual_updateYHat(p)
    return d, i - 1


def vbmf_dual(Y, p_in, niter, **kw):
    """src/vbmf_dual.jl:538-549 -> (params, d)."""
    p = _copy.deepcopy(p_in)
    d, _ = vbmf_dual_run(Y, p, niter, **kw)
    return p, d


def dual_lowerBound(Y, p):
    """src/vbmf_dual.jl:556-599."""
    Lb, GB = _lb_common(Y, p)
    eln0 = float(np.sum(gammaELn(p.alpha0 * np.ones(p.beta0.shape), p.beta0)))
    eln1 = float(np.sum(gammaELn(p.alpha1 * np.ones(p.beta1.shape), p.beta1)))
    eln_d = float(np.sum(gammaELn(p.gamma * np.ones(p.delta.shape), p.delta)))
    N0, N1 = p.M * p.H0, p.M * p.H1
    Lb += -p.MH / 2 * ln2pi + 0.5 * eln0
    Lb += 0.5 * eln1
    Lb += -(0.5 * float(p.CA @ (p.ATVecHat ** 2 + p.diagSigmaATVec)))
    Lb += -p.L * p.H / 2 * ln2pi
    Lb += p.L / 2 * eln_d
    Lb += -0.5 * traceXTY(np.diag(p.CB), GB)
    Lb += p.eta0 * math.log(p.zeta0) - gammaln(p.eta0)
    Lb += (p.eta0 - 1) * gammaELn(p.eta, p.zeta) - p.zeta0 * p.sigmaHat
    Lb += N0 * (p.alpha00 * math.log(p.beta00) - gammaln(p.alpha00))
    Lb += (p.alpha00 - 1) * eln0
    Lb += -p.beta00 * float(np.sum(p.CA0))
    Lb += N1 * (p.alpha01 * math.log(p.beta01) - gammaln(p.alpha01))
    Lb += (p.alpha01 - 1) * eln1
    Lb += -p.beta01 * float(np.sum(p.CA1))
    Lb += p.H * (p.gamma0 * math.log(p.delta0) - gammaln(p.gamma0))
    Lb += (p.gamma0 - 1) * eln_d
    Lb += -p.gamma0 * float(np.sum(p.CB))
    Lb += normalEntropy_diag(p.diagSigmaATVec)
    Lb += normalEntropy_kronSigmaB(p.SigmaB, p.L)
    Lb += gammaEntropy(p.eta, p.zeta)
    Lb += float(np.sum(gammaEntropy(p.alpha0 * np.ones(p.beta0.shape), p.beta0)))
    Lb += float(np.sum(gammaEntropy(p.alpha1 * np.ones(p.beta1.shape), p.beta1)))
    Lb += float(np.sum(gammaEntropy(p.gamma * np.ones(p.delta.shape), p.delta)))
    return float(Lb)


def dual_lowerBoundTrimmed(Y, p_in, trim=1e-1):
    """src/vbmf_dual.jl:606-617."""
    return dual_lowerBound(Y, _trim(p_in, trim))


# ----------------------------------------------------------------------------- vbmf_trial (src/vbmf_trial.jl, unexported)
def _interleave3(p, v1, v2, v3):
    """[CA1 block m ; CA2 block m] for m <= M0, [CA1 block m ; CA3 block m-M0] after (src/vbmf_trial.jl:176-183,383-390)."""
    M, H0, H1, M0 = p.M, p.H0, p.H1, p.M0
    out = np.empty((M, H0 + H1))
    out[:, :H0] = v1.reshape(M, H0)
    out[:M0, H0:] = v2.reshape(M0, H1)
    out[M0:, H0:] = v3.reshape(M - M0, H1)
    return out.reshape(M * (H0 + H1))


def vbmf_trial_init(Y, H, H0, M0, ca=1.0, alpha0=1e-10, beta0=1e-10, cb=1.0, gamma0=1e-10, delta0=1e-10, sigma=1.0,
                    eta0=1e-10, zeta0=1e-10, rng=None, AHat=None, BHat=None):
    """src/vbmf_trial.jl:139-227."""
    if H < H0:
        raise ValueError("H must be at least H0!")
    p = vbmf_dual_init(Y, H, H0, ca=ca, alpha0=alpha0, beta0=beta0, cb=cb, gamma0=gamma0, delta0=delta0, sigma=sigma,
                       eta0=eta0, zeta0=zeta0, rng=rng, AHat=AHat, BHat=BHat)
    for f in ("A0Hat", "A1Hat", "CA0", "CA1", "alpha00", "alpha01", "beta00", "beta01", "alpha0", "alpha1", "beta0", "beta1"):
        delattr(p, f)
    p.kind = "trial"
    M = p.M
    p.M0, p.M1 = int(M0), M - int(M0)
    p.A1Hat = p.AHat[:, :H0].copy()
    p.A2Hat = p.AHat[:M0, H0:].copy()
    p.A3Hat = p.AHat[M0:, H0:].copy()
    p.CA1, p.CA2, p.CA3 = ca * np.ones(M * H0), ca * np.ones(p.M0 * p.H1), ca * np.ones(p.M1 * p.H1)
    p.CA = _interleave3(p, p.CA1, p.CA2, p.CA3)
    p.alpha01 = p.alpha02 = p.alpha03 = alpha0
    p.beta01 = p.beta02 = p.beta03 = beta0
    p.alpha1 = p.alpha2 = p.alpha3 = alpha0 + 0.5
    p.beta1, p.beta2, p.beta3 = beta0 * np.ones(M * H0), beta0 * np.ones(p.M0 * p.H1), beta0 * np.ones(p.M1 * p.H1)
    p.alpha = np.array([p.alpha1, p.alpha2, p.alpha3])
    p.beta = _interleave3(p, p.beta1, p.beta2, p.beta3)
    return p


def trial_updateA(Y, p, full_cov=False, diag_var=False, literal=False):
    """src/vbmf_trial.jl:250-320."""
    _sparse_updateA_core(Y, p, full_cov, diag_var, literal)
    p.AHat = p.ATVecHat.reshape(p.M, p.H).copy()
    p.A1Hat = p.AHat[:, :p.H0].copy()
    p.A2Hat = p.AHat[:p.M0, p.H0:].copy()
    p.A3Hat = p.AHat[p.M0:, p.H0:].copy()


def trial_updateCA(p):
    """src/vbmf_trial.jl:357-400."""
    M, H, H0, H1, M0, M1 = p.M, p.H, p.H0, p.H1, p.M0, p.M1
    p.alpha1, p.alpha2, p.alpha3 = p.alpha01 + 0.5, p.alpha02 + 0.5, p.alpha03 + 0.5
    dS = p.diagSigmaATVec.reshape(M, H)
    p.beta1 = p.beta01 * np.ones(M * H0) + 0.5 * (p.A1Hat * p.A1Hat + dS[:, :H0]).reshape(M * H0)
    p.beta2 = p.beta02 * np.ones(M0 * H1) + 0.5 * (p.A2Hat * p.A2Hat).reshape(M0 * H1) + 0.5 * dS[:M0, H0:].reshape(M0 * H1)
    p.beta3 = p.beta03 * np.ones(M1 * H1) + 0.5 * (p.A3Hat * p.A3Hat + dS[M0:, H0:]).reshape(M1 * H1)
    p.CA1 = p.alpha1 * np.ones(M * H0) / p.beta1
    p.CA2 = p.alpha2 * np.ones(M0 * H1) / p.beta2
    p.CA3 = p.alpha3 * np.ones(M1 * H1) / p.beta3
    p.CA = _interleave3(p, p.CA1, p.CA2, p.CA3)
    p.alpha = np.array([p.alpha1, p.alpha2, p.alpha3])
    p.beta = _interleave3(p, p.beta1, p.beta2, p.beta3)


def trial_update_priors(p):
    """updateAlpha01!/02!/03! then updateBeta01!/02!/03! (src/vbmf_trial.jl:442-507, order :566-573)."""
    N = (p.M * p.H0, p.M0 * p.H1, p.M1 * p.H1)
    p.alpha01 = _update_alpha0x(N[0], p.beta01, p.alpha1, p.beta1, p.alpha01)
    p.alpha02 = _update_alpha0x(N[1], p.beta02, p.alpha2, p.beta2, p.alpha02)
    p.alpha03 = _update_alpha0x(N[2], p.beta03, p.alpha3, p.beta3, p.alpha03)
    with np.errstate(divide="ignore", invalid="ignore"):
        p.beta01 = float(np.float64(N[0] * p.alpha01) / np.sum(p.CA1))
        p.beta02 = float(np.float64(N[1] * p.alpha02) / np.sum(p.CA2))
        p.beta03 = float(np.float64(N[2] * p.alpha03) / np.sum(p.CA3))


def vbmf_trial_run(Y, p, niter, eps=1e-6, diag_var=False, full_cov=False, est_priors=True, est_cb=True, norm="spectral", trace=None):
    """`vbmf_trial!` src/vbmf_trial.jl:528-604.  Returns (d, iterations_done)."""
    old = p.BHat.copy()
    d = eps + 1.0
    i = 1
    while i <= niter and d > eps:
        trial_updateA(Y, p, full_cov=full_cov, diag_var=diag_var)
        sparse_updateB(Y, p, diag_var=diag_var)          # src/vbmf_trial.jl:327-341 == sparse
        trial_updateCA(p)
        if est_cb:
            sparse_updateCB(p)                           # :407-412
        sparse_updateSigma(Y, p, diag_var=diag_var)      # :419-435
        if est_priors:
            trial_update_priors(p)
        if trace is not None:
            trace(p, i)
        d = delta(p.BHat, old, norm)
        old = p.BHat.copy()
        i += 1
    sparse_updateYHat(p)
    return d, i - 1


def trial_lowerBound(Y, p):
    """src/vbmf_trial.jl:630-681."""
    Lb, GB = _lb_common(Y, p)
    groups = ((p.alpha1, p.beta1, p.alpha01, p.beta01, p.CA1, p.M * p.H0), (p.alpha2, p.beta2, p.alpha02, p.beta02, p.CA2, p.M0 * p.H1),
              (p.alpha3, p.beta3, p.alpha03, p.beta03, p.CA3, p.M1 * p.H1))
    eln = [float(np.sum(gammaELn(a * np.ones(b.shape), b))) for a, b, _, _, _, _ in groups]
    eln_d = float(np.sum(gammaELn(p.gamma * np.ones(p.delta.shape), p.delta)))
    Lb += -p.MH / 2 * ln2pi + 0.5 * eln[0]
    Lb += 0.5 * eln[1]
    Lb += 0.5 * eln[2]
    Lb += -(0.5 * float(p.CA @ (p.ATVecHat ** 2 + p.diagSigmaATVec)))
    Lb += -p.L * p.H / 2 * ln2pi
    Lb += p.L / 2 * eln_d
    Lb += -0.5 * traceXTY(np.diag(p.CB), GB)
    Lb += p.eta0 * math.log(p.zeta0) - gammaln(p.eta0)
    Lb += (p.eta0 - 1) * gammaELn(p.eta, p.zeta) - p.zeta0 * p.sigmaHat
    for (a, b, a0, b0, ca, N), e in zip(groups, eln):
        Lb += N * (a0 * math.log(b0) - gammaln(a0))
        Lb += (a0 - 1) * e
        Lb += -b0 * float(np.sum(ca))
    Lb += p.H * (p.gamma0 * math.log(p.delta0) - gammaln(p.gamma0))
    Lb += (p.gamma0 - 1) * eln_d
    Lb += -p.gamma0 * float(np.sum(p.CB))
    Lb += normalEntropy_diag(p.diagSigmaATVec)
    Lb += normalEntropy_kronSigmaB(p.SigmaB, p.L)
    Lb += gammaEntropy(p.eta, p.zeta)
    for a, b, _, _, _, _ in groups:
        Lb += float(np.sum(gammaEntropy(a * np.ones(b.shape), b)))
    Lb += float(np.sum(gammaEntropy(p.gamma * np.ones(p.delta.shape), p.delta)))
    return float(Lb)


def trial_lowerBoundTrimmed(Y, p_in, trim=1e-1):
    """src/vbmf_trial.jl:687-697."""
    return trial_lowerBound(Y, _trim(p_in, trim))


# ----------------------------------------------------------------------------- vbls! (examples/mil_util.jl:179-203)
def vbls(Y, p, niter, diag_var=False, full_cov=False):
    """A-only VB with B fixed (the MIL classification call pattern)."""
    for _ in range(niter):
        if p.kind == "dense":
            dense_updateA(Y, p)
            dense_updateCA(p)
            dense_updateSigma2(Y, p)
        elif p.kind == "sparse":
            sparse_updateA(Y, p, full_cov=full_cov, diag_var=diag_var)
            sparse_updateCA(p)
            sparse_updateSigma(Y, p, diag_var=diag_var)
        elif p.kind == "trial":
            trial_updateA(Y, p, full_cov=full_cov, diag_var=diag_var)
            trial_updateCA(p)
            sparse_updateSigma(Y, p, diag_var=diag_var)
        else:
            dual_updateA(Y, p, full_cov=full_cov, diag_var=diag_var)
            dual_updateCA(p)
            dual_updateSigma(Y, p, diag_var=diag_var)
    p.YHat = p.BHat @ p.AHat.T
    return p.AHat
