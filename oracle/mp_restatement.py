"""Second, independent restatement of the un-pinned branches, in 40-digit arithmetic (mpmath).

TEST INFRASTRUCTURE ONLY (same rule as vbmf_oracle.py: imported by tests/ and nothing else).

SURVEY section 8(c) lists the branches of the reference that none of its own golden logs exercise (`full_cov=false` with the
`repeat(inner=M-1)` tail, `diag_var=true`, the label mask, `vbmf_dual`'s grouped ARD and hyper-prior root, `lowerBound`).
For those the NumPy oracle IS the specification, so this file transcribes the same reference lines a second time in a
different style: the LITERAL formulae - the (MH)x(MH) Kronecker precision and its dense inverse, `diagm(sigmaVecHat)` as a
dense L x L matrix, `repeat(...; inner=M-1)` as an actual repeated list, the determinant of the (LH)x(LH) `kron(SigmaB, I)` -
evaluated with mpmath matrices at 40 digits, so rounding plays no part and an algebraic slip in either restatement
shows up as a disagreement far above 1e-12.  Sizes are tiny (M*H <= 16); everything is O((MH)^3) as in the reference.

The state is a plain dict whose keys are the reference's field names; vectors are Python lists of mpf, matrices mp.matrix.
"""
import mpmath as mp

mp.mp.dps = 40
LN2PI = mp.log(2 * mp.pi)


def M_(a):
    """numpy 2-D array -> mp.matrix (empty shapes allowed)"""
    import numpy as np
    a = np.asarray(a)
    if a.size == 0:
        return mp.matrix(int(a.shape[0]), int(a.shape[1]))
    return mp.matrix([[mp.mpf(float(x)) for x in row] for row in a])


def V_(a):
    return [mp.mpf(float(x)) for x in a]


def to_float(x):
    if isinstance(x, mp.matrix):
        return [[float(x[i, j]) for j in range(x.cols)] for i in range(x.rows)]
    if isinstance(x, list):
        return [float(v) for v in x]
    return float(x)


def eye(n):
    return mp.eye(n)


def diagm(v):
    n = len(v)
    D = mp.zeros(n, n)
    for i in range(n):
        D[i, i] = v[i]
    return D


def kron(A, B):
    K = mp.zeros(A.rows * B.rows, A.cols * B.cols)
    for i in range(A.rows):
        for j in range(A.cols):
            for k in range(B.rows):
                for l in range(B.cols):
                    K[i * B.rows + k, j * B.cols + l] = A[i, j] * B[k, l]
    return K


def vec_colmajor(X):
    """Julia reshape(X, length) / vec(X): column-major."""
    return [X[i, j] for j in range(X.cols) for i in range(X.rows)]


def reshape_colmajor(v, r, c):
    X = mp.zeros(r, c)
    for j in range(c):
        for i in range(r):
            X[i, j] = v[j * r + i]
    return X


def mean(v):
    return mp.fsum(v) / len(v)


def repeat_inner(v, inner):
    """Julia repeat(v, inner = k): every element k times in a row."""
    out = []
    for x in v:
        out.extend([x] * inner)
    return out


def trace_xty(X, Y):
    return mp.fsum(X[i, j] * Y[i, j] for i in range(X.rows) for j in range(X.cols))


def gammaELn(a, b):            # src/util.jl:153-155
    return mp.digamma(a) - mp.log(b)


def gammaEntropy(a, b):        # src/util.jl:144-146 (+log b as written)
    return a + mp.log(b) + mp.loggamma(a) + (1 - a) * mp.digamma(a)


EPS0 = mp.mpf(2) ** -1074      # eps(0.0)


def normalEntropy_mat(S):      # src/util.jl:113-125
    d = mp.det(S)
    if d < EPS0:
        d = EPS0
    m = S.rows
    return mp.mpf(m) / 2 + mp.mpf(m) / 2 * LN2PI + mp.log(d) / 2


def normalEntropy_diag(v):     # src/util.jl:133-137
    n = len(v)
    return mp.mpf(n) / 2 + mp.mpf(n) / 2 * LN2PI + mp.fsum(mp.log(x) for x in v) / 2


# ------------------------------------------------------------------------------------------------ src/vbmf_sparse.jl
def updateA(Y, p, full_cov=False, diag_var=False, mask=True):
    """src/vbmf_sparse.jl:176-247 (and src/vbmf_dual.jl:216-287 with mask=False)."""
    L, M, H = p["L"], p["M"], p["H"]
    B = p["BHat"]
    if full_cov:
        if diag_var:
            G = B.T * diagm(p["sigmaVecHat"]) * B + L * mean(p["sigmaVecHat"]) * p["SigmaB"]
            inv_sigma = kron(eye(M), G) + diagm(p["CA"])
        else:
            inv_sigma = p["sigmaHat"] * kron(eye(M), B.T * B + L * p["SigmaB"]) + diagm(p["CA"])
        sigma = inv_sigma ** -1
        p["diagSigmaATVec"] = [sigma[i, i] for i in range(M * H)]
        if diag_var:
            rhs = vec_colmajor(B.T * diagm(p["sigmaVecHat"]) * Y)
            at = sigma * mp.matrix(rhs)
        else:
            rhs = vec_colmajor(B.T * Y)
            at = p["sigmaHat"] * sigma * mp.matrix(rhs)
        p["ATVecHat"] = [at[i] for i in range(M * H)]
        SA = mp.zeros(H, H)
        for m in range(M):
            for a in range(H):
                for b in range(H):
                    SA[a, b] += sigma[m * H + a, m * H + b]
        p["SigmaA"] = SA
        p["SigmaATVec"] = sigma
    else:
        d = list(p["diagSigmaATVec"])
        for h in range(H):
            col = [B[l, h] for l in range(L)]
            if diag_var:
                d[h] = mp.fsum((col[l] * p["sigmaVecHat"][l]) ** 2 for l in range(L)) + L * mean(p["sigmaVecHat"]) * p["SigmaB"][h, h]
            else:
                d[h] = p["sigmaHat"] * mp.fsum(x * x for x in col) + L * p["SigmaB"][h, h]
        d[H:] = repeat_inner(d[:H], M - 1)
        d = [d[i] + p["CA"][i] for i in range(M * H)]
        d = [1 / x for x in d]
        p["diagSigmaATVec"] = d
        if diag_var:
            rhs = vec_colmajor(B.T * diagm(p["sigmaVecHat"]) * Y)
            p["ATVecHat"] = [d[i] * rhs[i] for i in range(M * H)]
        else:
            rhs = vec_colmajor(B.T * Y)
            p["ATVecHat"] = [p["sigmaHat"] * d[i] * rhs[i] for i in range(M * H)]
        SA = mp.zeros(H, H)
        for m in range(M):
            for h in range(H):
                SA[h, h] += d[m * H + h]
        p["SigmaA"] = SA
    A = reshape_colmajor(p["ATVecHat"], H, M).T
    if mask and p.get("H1", 0) > 0:
        for lab in p.get("labels", []):           # 1-based row labels
            for h in range(H - p["H1"], H):
                A[lab - 1, h] = 0
    p["AHat"] = A
    p["ATVecHat"] = vec_colmajor(A.T)
    if not mask:                                   # dual: A0Hat / A1Hat views, src/vbmf_dual.jl:284-286
        H0 = p["H0"]
        p["A0Hat"] = A[:, 0:H0]
        p["A1Hat"] = A[:, H0:H]


def updateB(Y, p, diag_var=False):
    """src/vbmf_sparse.jl:254-268."""
    A = p["AHat"]
    if diag_var:
        S = diagm(p["CB"]) + mean(p["sigmaVecHat"]) * (A.T * A + p["SigmaA"])
        S = S ** -1
        p["SigmaB"] = S
        p["BHat"] = diagm(p["sigmaVecHat"]) * Y * A * S
    else:
        S = diagm(p["CB"]) + p["sigmaHat"] * (A.T * A + p["SigmaA"])
        S = S ** -1
        p["SigmaB"] = S
        p["BHat"] = p["sigmaHat"] * Y * A * S


def updateCA(p):
    """src/vbmf_sparse.jl:284-288."""
    n = p["M"] * p["H"]
    p["beta"] = [p["beta0"] + (p["ATVecHat"][i] ** 2 + p["diagSigmaATVec"][i]) / 2 for i in range(n)]
    p["CA"] = [p["alpha"] / b for b in p["beta"]]


def updateCB(p):
    """src/vbmf_sparse.jl:295-300."""
    B = p["BHat"]
    for h in range(p["H"]):
        p["delta"][h] = p["delta0"] + mp.fsum(B[l, h] ** 2 for l in range(p["L"])) / 2 + p["SigmaB"][h, h] / 2
        p["CB"][h] = p["gamma"] / p["delta"][h]


def updateSigma(Y, p, diag_var=False):
    """src/vbmf_sparse.jl:307-323."""
    A, B, L = p["AHat"], p["BHat"], p["L"]
    GA = A.T * A + p["SigmaA"]
    if diag_var:
        for l in range(L):
            yl = Y[l, :].T
            bl = B[l, :].T
            ab = A * bl
            p["zetaVec"][l] = (p["zeta0"] + mp.fsum(x * x for x in yl) / 2 - mp.fsum(yl[i] * ab[i] for i in range(p["M"]))
                               + trace_xty(GA, bl * bl.T + p["SigmaB"]) / 2)
            p["sigmaVecHat"][l] = p["etaVec"][l] / p["zetaVec"][l]
    else:
        p["zeta"] = p["zeta0"] + p["trYTY"] / 2 - trace_xty(B, Y * A) + trace_xty(GA, B.T * B + L * p["SigmaB"]) / 2
        p["sigmaHat"] = p["eta"] / p["zeta"]


def lowerBound(Y, p):
    """src/vbmf_sparse.jl:435-471, term by term, with the literal (LH)x(LH) Kronecker determinant."""
    L, M, H, MH = p["L"], p["M"], p["H"], p["MH"]
    A, B = p["AHat"], p["BHat"]
    Lb = mp.mpf(0)
    Lb += -mp.mpf(L * M) / 2 * LN2PI + mp.mpf(L * M) / 2 * gammaELn(p["eta"], p["zeta"])
    GB = B.T * B + L * p["SigmaB"]
    Lb += -p["sigmaHat"] / 2 * (p["trYTY"] - 2 * trace_xty(B, Y * A) + trace_xty(A.T * A + p["SigmaA"], GB))
    eln_b = mp.fsum(gammaELn(p["alpha"], b) for b in p["beta"])
    Lb += -mp.mpf(MH) / 2 * LN2PI + eln_b / 2
    Lb += -mp.fsum(p["CA"][i] * (p["ATVecHat"][i] ** 2 + p["diagSigmaATVec"][i]) for i in range(len(p["CA"]))) / 2
    Lb += -mp.mpf(L * H) / 2 * LN2PI
    eln_d = mp.fsum(gammaELn(p["gamma"], d) for d in p["delta"])
    Lb += mp.mpf(L) / 2 * eln_d
    Lb += -trace_xty(diagm(p["CB"]), GB) / 2
    Lb += p["eta0"] * mp.log(p["zeta0"]) - mp.loggamma(p["eta0"])
    Lb += (p["eta0"] - 1) * gammaELn(p["eta"], p["zeta"]) - p["zeta0"] * p["sigmaHat"]
    Lb += MH * (p["alpha0"] * mp.log(p["beta0"]) - mp.loggamma(p["alpha0"]))
    Lb += (p["alpha0"] - 1) * eln_b
    Lb += -p["beta0"] * mp.fsum(p["CA"])
    Lb += H * (p["gamma0"] * mp.log(p["delta0"]) - mp.loggamma(p["gamma0"]))
    Lb += (p["gamma0"] - 1) * eln_d
    Lb += -p["gamma0"] * mp.fsum(p["CB"])
    Lb += normalEntropy_diag(p["diagSigmaATVec"])
    Lb += normalEntropy_mat(kron(p["SigmaB"], eye(L)))
    Lb += gammaEntropy(p["eta"], p["zeta"])
    Lb += mp.fsum(gammaEntropy(p["alpha"], b) for b in p["beta"])
    Lb += mp.fsum(gammaEntropy(p["gamma"], d) for d in p["delta"])
    return Lb


# ------------------------------------------------------------------------------------------------ src/vbmf_dual.jl
def dual_updateCA(p):
    """src/vbmf_dual.jl:322-351."""
    M, H, H0 = p["M"], p["H"], p["H0"]
    H1 = H - H0
    p["alpha0"] = p["alpha00"] + mp.mpf(1) / 2
    p["alpha1"] = p["alpha01"] + mp.mpf(1) / 2
    dS = reshape_colmajor(p["diagSigmaATVec"], H, M)
    dS0, dS1 = dS[0:H0, :], dS[H0:H, :]
    A0t, A1t = p["A0Hat"].T, p["A1Hat"].T            # H0 x M, H1 x M
    t0 = vec_colmajor(mp.matrix([[A0t[i, j] ** 2 + dS0[i, j] for j in range(M)] for i in range(H0)])) if H0 else []
    t1 = vec_colmajor(mp.matrix([[A1t[i, j] ** 2 + dS1[i, j] for j in range(M)] for i in range(H1)])) if H1 else []
    p["beta0"] = [p["beta00"] + x / 2 for x in t0]
    p["beta1"] = [p["beta01"] + x / 2 for x in t1]
    p["CA0"] = [p["alpha0"] / b for b in p["beta0"]]
    p["CA1"] = [p["alpha1"] / b for b in p["beta1"]]
    CA, beta = [], []
    for m in range(M):
        CA += p["CA0"][m * H0:(m + 1) * H0] + p["CA1"][m * H1:(m + 1) * H1]
        beta += p["beta0"][m * H0:(m + 1) * H0] + p["beta1"][m * H1:(m + 1) * H1]
    p["CA"], p["beta"] = CA, beta
    p["alpha"] = [p["alpha0"], p["alpha1"]]


def dual_lowerBound(Y, p):
    """src/vbmf_dual.jl:556-599, term by term (two CA groups), with the literal (LH)x(LH) Kronecker determinant."""
    L, M, H, H0, MH = p["L"], p["M"], p["H"], p["H0"], p["MH"]
    H1 = H - H0
    A, B = p["AHat"], p["BHat"]
    GB = B.T * B + L * p["SigmaB"]
    e0 = mp.fsum(gammaELn(p["alpha0"], b) for b in p["beta0"])
    e1 = mp.fsum(gammaELn(p["alpha1"], b) for b in p["beta1"])
    ed = mp.fsum(gammaELn(p["gamma"], d) for d in p["delta"])
    Lb = mp.mpf(0)
    Lb += -mp.mpf(L * M) / 2 * LN2PI + mp.mpf(L * M) / 2 * gammaELn(p["eta"], p["zeta"])
    Lb += -p["sigmaHat"] / 2 * (p["trYTY"] - 2 * trace_xty(B, Y * A) + trace_xty(A.T * A + p["SigmaA"], GB))
    Lb += -mp.mpf(MH) / 2 * LN2PI + e0 / 2
    Lb += e1 / 2
    Lb += -mp.fsum(p["CA"][i] * (p["ATVecHat"][i] ** 2 + p["diagSigmaATVec"][i]) for i in range(len(p["CA"]))) / 2
    Lb += -mp.mpf(L * H) / 2 * LN2PI
    Lb += mp.mpf(L) / 2 * ed
    Lb += -trace_xty(diagm(p["CB"]), GB) / 2
    Lb += p["eta0"] * mp.log(p["zeta0"]) - mp.loggamma(p["eta0"])
    Lb += (p["eta0"] - 1) * gammaELn(p["eta"], p["zeta"]) - p["zeta0"] * p["sigmaHat"]
    Lb += M * H0 * (p["alpha00"] * mp.log(p["beta00"]) - mp.loggamma(p["alpha00"]))
    Lb += (p["alpha00"] - 1) * e0
    Lb += -p["beta00"] * mp.fsum(p["CA0"])
    Lb += M * H1 * (p["alpha01"] * mp.log(p["beta01"]) - mp.loggamma(p["alpha01"]))
    Lb += (p["alpha01"] - 1) * e1
    Lb += -p["beta01"] * mp.fsum(p["CA1"])
    Lb += H * (p["gamma0"] * mp.log(p["delta0"]) - mp.loggamma(p["gamma0"]))
    Lb += (p["gamma0"] - 1) * ed
    Lb += -p["gamma0"] * mp.fsum(p["CB"])
    Lb += normalEntropy_diag(p["diagSigmaATVec"])
    Lb += normalEntropy_mat(kron(p["SigmaB"], eye(L)))
    Lb += gammaEntropy(p["eta"], p["zeta"])
    Lb += mp.fsum(gammaEntropy(p["alpha0"], b) for b in p["beta0"])
    Lb += mp.fsum(gammaEntropy(p["alpha1"], b) for b in p["beta1"])
    Lb += mp.fsum(gammaEntropy(p["gamma"], d) for d in p["delta"])
    return Lb


def _root(f, a=mp.mpf("1e-10"), b=mp.mpf("1e10")):
    """Exact root of a monotone f on [a, b] (what a bracketing solver run to floating-point resolution returns, to within
    one ulp); None when there is no sign change - the reference's `try ... end` then keeps the old value."""
    fa, fb = f(a), f(b)
    if fa * fb > 0:
        return None
    return mp.findroot(f, (a, b), solver="anderson", tol=mp.mpf(10) ** -60, maxsteps=2000)


def dual_update_priors(p):
    """src/vbmf_dual.jl:393-434 in the order of the main loop (:493-498): alpha00, alpha01, beta00, beta01."""
    M, H0 = p["M"], p["H0"]
    H1 = p["H"] - H0
    for g, (N, a0x, b0x, ag, bg) in enumerate(((M * H0, "alpha00", "beta00", "alpha0", "beta0"), (M * H1, "alpha01", "beta01", "alpha1", "beta1"))):
        s = mp.fsum(gammaELn(p[ag], b) for b in p[bg])
        r = _root(lambda x: N * mp.log(p[b0x]) - N * mp.digamma(x) + s) if N > 0 else None
        if r is not None:
            p[a0x] = r
    if M * H0 > 0:
        p["beta00"] = M * H0 * p["alpha00"] / mp.fsum(p["CA0"])
    if M * H1 > 0:
        p["beta01"] = M * H1 * p["alpha01"] / mp.fsum(p["CA1"])


# ------------------------------------------------------------------------------------------------ src/vbmf_trial.jl
def trial_split(p):
    """src/vbmf_trial.jl:316-319: the three group views of AHat."""
    A, H, H0, M, M0 = p["AHat"], p["H"], p["H0"], p["M"], p["M0"]
    p["A1Hat"] = A[:, 0:H0]
    p["A2Hat"] = A[0:M0, H0:H] if M0 > 0 and H > H0 else mp.matrix(M0, H - H0)
    p["A3Hat"] = A[M0:M, H0:H] if M > M0 and H > H0 else mp.matrix(M - M0, H - H0)


def _sq_plus(At, dS):
    """vec(At .* At + dS), column-major, for mp matrices of equal shape (possibly empty)."""
    if At.rows == 0 or At.cols == 0:
        return []
    return vec_colmajor(mp.matrix([[At[i, j] ** 2 + dS[i, j] for j in range(At.cols)] for i in range(At.rows)]))


def trial_updateCA(p):
    """src/vbmf_trial.jl:357-400."""
    M, H, H0, M0 = p["M"], p["H"], p["H0"], p["M0"]
    H1, M1 = H - H0, M - M0
    half = mp.mpf(1) / 2
    p["alpha1"], p["alpha2"], p["alpha3"] = p["alpha01"] + half, p["alpha02"] + half, p["alpha03"] + half
    dS = reshape_colmajor(p["diagSigmaATVec"], H, M)
    z = lambda r, c: mp.matrix(r, c)
    dS1 = dS[0:H0, :] if H0 else z(0, M)
    dS2 = dS[H0:H, 0:M0] if (H1 and M0) else z(H1, M0)
    dS3 = dS[H0:H, M0:M] if (H1 and M1) else z(H1, M1)
    p["beta1"] = [p["beta01"] + x / 2 for x in _sq_plus(p["A1Hat"].T if H0 else z(0, M), dS1)]
    p["beta2"] = [p["beta02"] + x / 2 for x in _sq_plus(p["A2Hat"].T if (H1 and M0) else z(H1, M0), dS2)]
    p["beta3"] = [p["beta03"] + x / 2 for x in _sq_plus(p["A3Hat"].T if (H1 and M1) else z(H1, M1), dS3)]
    p["CA1"] = [p["alpha1"] / b for b in p["beta1"]]
    p["CA2"] = [p["alpha2"] / b for b in p["beta2"]]
    p["CA3"] = [p["alpha3"] / b for b in p["beta3"]]
    CA, beta = [], []
    for m in range(M0):
        CA += p["CA1"][m * H0:(m + 1) * H0] + p["CA2"][m * H1:(m + 1) * H1]
        beta += p["beta1"][m * H0:(m + 1) * H0] + p["beta2"][m * H1:(m + 1) * H1]
    for m in range(M0, M):
        CA += p["CA1"][m * H0:(m + 1) * H0] + p["CA3"][(m - M0) * H1:(m - M0 + 1) * H1]
        beta += p["beta1"][m * H0:(m + 1) * H0] + p["beta3"][(m - M0) * H1:(m - M0 + 1) * H1]
    p["CA"], p["beta"] = CA, beta
    p["alpha"] = [p["alpha1"], p["alpha2"], p["alpha3"]]


def trial_update_priors(p):
    """src/vbmf_trial.jl:442-507 in the loop's order (:565-570): alpha01, alpha02, alpha03, beta01, beta02, beta03."""
    M, H0, M0 = p["M"], p["H0"], p["M0"]
    H1, M1 = p["H"] - H0, M - M0
    groups = ((M * H0, "alpha01", "beta01", "alpha1", "beta1", "CA1"), (M0 * H1, "alpha02", "beta02", "alpha2", "beta2", "CA2"),
              (M1 * H1, "alpha03", "beta03", "alpha3", "beta3", "CA3"))
    for N, a0x, b0x, ag, bg, _ in groups:
        if N == 0:
            continue                                   # f == 0 everywhere: no bracket, the `try` keeps the old value
        s = mp.fsum(gammaELn(p[ag], b) for b in p[bg])
        r = _root(lambda x: N * mp.log(p[b0x]) - N * mp.digamma(x) + s)
        if r is not None:
            p[a0x] = r
    for N, a0x, b0x, _, _, cag in groups:
        if N > 0:
            p[b0x] = N * p[a0x] / mp.fsum(p[cag])


def from_oracle(p, Y):
    """SimpleNamespace state of oracle/vbmf_oracle.py -> (mp Y, mp state dict)."""
    import numpy as np
    d = {}
    for k, v in vars(p).items():
        if v is None or k in ("kind", "SigmaATVec_blocks", "YHat"):
            continue
        if isinstance(v, (int, np.integer)):
            d[k] = int(v)
        elif isinstance(v, (float, np.floating)):
            d[k] = mp.mpf(float(v))
        else:
            a = np.asarray(v)
            if a.dtype.kind == "i":
                d[k] = [int(x) for x in a]
            elif a.ndim == 1:
                d[k] = V_(a)
            else:
                d[k] = M_(a)
    return M_(Y), d
