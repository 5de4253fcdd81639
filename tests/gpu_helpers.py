"""Helpers for the GPU parity tests: oracle params <-> vbmf_b200 params conversion."""
import numpy as np

import vbmf_b200_loader

vb = vbmf_b200_loader.load()


def to_gpu_params(p):
    """Oracle SimpleNamespace -> vbmf_b200 parameter object (deep copies, Fortran order)."""
    if p.kind == "dense":
        q = vb.vbmf_parameters()
        for f in ("L", "M", "H", "H1"):
            setattr(q, f, int(getattr(p, f)))
        q.labels = np.asarray(p.labels, dtype=np.int64).copy()
        for f in ("AHat", "BHat", "SigmaA", "SigmaB", "CA", "CB", "invCA", "invCB"):
            setattr(q, f, np.asfortranarray(np.array(getattr(p, f), dtype=np.float64)))
        q.sigma2 = float(p.sigma2)
        q.YHat = None
        return q
    q = {"sparse": vb.vbmf_sparse_parameters, "dual": vb.vbmf_dual_parameters, "trial": vb.vbmf_trial_parameters}[p.kind]()
    for f in ("L", "M", "H", "MH", "H1"):
        setattr(q, f, int(getattr(p, f)))
    q.labels = np.asarray(getattr(p, "labels", []), dtype=np.int64).copy()
    for f in ("AHat", "BHat", "SigmaA", "SigmaB"):
        setattr(q, f, np.asfortranarray(np.array(getattr(p, f), dtype=np.float64)))
    for f in ("ATVecHat", "diagSigmaATVec", "CA", "beta", "CB", "delta", "sigmaVecHat", "etaVec", "zetaVec"):
        setattr(q, f, np.array(getattr(p, f), dtype=np.float64).copy())
    for f in ("gamma0", "delta0", "gamma", "sigmaHat", "eta0", "zeta0", "eta", "zeta", "trYTY"):
        setattr(q, f, float(getattr(p, f)))
    q.SigmaATVec_blocks = None
    q.YHat = None
    if p.kind == "sparse":
        for f in ("alpha0", "beta0", "alpha"):
            setattr(q, f, float(getattr(p, f)))
    elif p.kind == "trial":
        q.H0, q.M0, q.M1 = int(p.H0), int(p.M0), int(p.M1)
        for f in ("A1Hat", "A2Hat", "A3Hat"):
            setattr(q, f, np.asfortranarray(np.array(getattr(p, f), dtype=np.float64)))
        for f in ("CA1", "CA2", "CA3", "beta1", "beta2", "beta3", "alpha"):
            setattr(q, f, np.array(getattr(p, f), dtype=np.float64).copy())
        for f in ("alpha01", "beta01", "alpha1", "alpha02", "beta02", "alpha2", "alpha03", "beta03", "alpha3"):
            setattr(q, f, float(getattr(p, f)))
    else:
        q.H0 = int(p.H0)
        for f in ("A0Hat", "A1Hat"):
            setattr(q, f, np.asfortranarray(np.array(getattr(p, f), dtype=np.float64)))
        for f in ("CA0", "CA1", "beta0", "beta1", "alpha"):
            setattr(q, f, np.array(getattr(p, f), dtype=np.float64).copy())
        for f in ("alpha00", "beta00", "alpha01", "beta01", "alpha0", "alpha1"):
            setattr(q, f, float(getattr(p, f)))
    return q


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return float("inf")
    if b.size == 0:
        return 0.0
    na, nb = np.isnan(a), np.isnan(b)
    if np.any(na != nb):
        return float("inf")
    if np.all(nb):
        return 0.0          # NaN where the reference formula gives NaN (e.g. 0/0 for an empty group)
    a, b = a[~nb], b[~nb]
    den = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b))) / den


_SP = ["AHat", "ATVecHat", "diagSigmaATVec", "SigmaA", "BHat", "SigmaB", "CA", "beta", "CB", "delta", "sigmaHat", "zeta", "sigmaVecHat", "zetaVec"]
FIELDS = {
    "trial": _SP + ["A1Hat", "A2Hat", "A3Hat", "CA1", "CA2", "CA3", "beta1", "beta2", "beta3", "alpha01", "beta01", "alpha02", "beta02",
                    "alpha03", "beta03", "alpha1", "alpha2", "alpha3"],
    "dense": ["AHat", "BHat", "SigmaA", "SigmaB", "CA", "CB", "invCA", "invCB", "sigma2"],
    "sparse": ["AHat", "ATVecHat", "diagSigmaATVec", "SigmaA", "BHat", "SigmaB", "CA", "beta", "CB", "delta", "sigmaHat",
               "zeta", "sigmaVecHat", "zetaVec"],
    "dual": ["AHat", "ATVecHat", "diagSigmaATVec", "SigmaA", "BHat", "SigmaB", "CA", "beta", "CB", "delta", "sigmaHat", "zeta",
             "sigmaVecHat", "zetaVec", "A0Hat", "A1Hat", "CA0", "CA1", "beta0", "beta1", "alpha00", "beta00", "alpha01",
             "beta01", "alpha0", "alpha1"],
}


# measured worst relative error per test (filled by compare, written to gpurun_out/parity_worst.json by tests/conftest.py)
WORST = {}
CURRENT = [""]


def record(name, err):
    d = WORST.setdefault(CURRENT[0], {})
    if not (d.get(name, -1.0) >= err):
        d[name] = float(err)


def compare(q, p, tol, fields=None):
    """max relative error per field between GPU params q and oracle params p; asserts all <= tol."""
    errs = {f: rel(getattr(q, f), getattr(p, f)) for f in (fields or FIELDS[p.kind])}
    for f, e in errs.items():
        record(f, e)
    bad = {f: e for f, e in errs.items() if not e <= tol}
    assert not bad, "fields above %.1e: %s (all: %s)" % (tol, bad, errs)
    return errs


def sensitivity(run, p, fields, scale=1e-14, seed=123):
    """How much the oracle's own trajectory moves under a 1e-14 relative perturbation of the initial factors: the floor
    below which no two Float64 implementations can agree after that many free-running iterations (ill-conditioned
    posterior precisions amplify rounding).  Returns max relative change over `fields`."""
    import copy
    rng = np.random.default_rng(seed)
    a, b = copy.deepcopy(p), copy.deepcopy(p)
    b.AHat = b.AHat * (1.0 + scale * rng.standard_normal(b.AHat.shape))
    b.BHat = b.BHat * (1.0 + scale * rng.standard_normal(b.BHat.shape))
    if hasattr(b, "ATVecHat"):
        b.ATVecHat = b.AHat.reshape(-1).copy()
    run(a)
    run(b)
    return max(rel(getattr(b, f), getattr(a, f)) for f in fields)
