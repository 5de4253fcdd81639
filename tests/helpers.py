"""Shared test helpers: golden loading and oracle state construction from the reference's logged slice 0."""
import os
from types import SimpleNamespace

import numpy as np

from oracle import vbmf_oracle as vo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: g[k] for k in g.files}


def jl(arr, t):
    """Julia matrix of logged slice t (HDF5 dims are reversed, see make_golden.py)."""
    a = arr[t]
    return np.ascontiguousarray(a.T) if a.ndim == 2 else a.copy()


def dense_state_from_golden(g, t=0):
    Y = np.ascontiguousarray(g["Y"].T)
    p = vo.vbmf_init(Y, int(g["log_H"][0]), AHat=jl(g["log_AHat"], t), BHat=jl(g["log_BHat"], t))
    for f in ("SigmaA", "SigmaB", "CA", "CB", "invCA", "invCB"):
        setattr(p, f, jl(g["log_" + f], t))
    p.sigma2 = float(g["log_sigma2"][t])
    p.H1 = int(g["log_H1"][0])
    return Y, p


def sparse_state_from_golden(g, t=0):
    Y = np.ascontiguousarray(g["Y"].T)
    H = int(g["log_H"][0])
    p = vo.vbmf_sparse_init(Y, H, AHat=jl(g["log_AHat"], t), BHat=jl(g["log_BHat"], t))
    for f in ("SigmaA", "SigmaB"):
        setattr(p, f, jl(g["log_" + f], t))
    for f in ("ATVecHat", "diagSigmaATVec", "CA", "beta", "CB", "delta", "sigmaVecHat", "etaVec", "zetaVec"):
        setattr(p, f, g["log_" + f][t].copy())
    for f in ("alpha0", "beta0", "alpha", "gamma0", "delta0", "gamma", "sigmaHat", "eta0", "zeta0", "eta", "zeta", "trYTY"):
        setattr(p, f, float(g["log_" + f][t]))
    p.H1 = int(g["log_H1"][0])
    return Y, p


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b))) / den


def synth(L, M, r, seed=0, noise=0.1):
    """Low-rank-plus-noise Y (L x M, Julia orientation) as in SURVEY 8(d)."""
    rng = np.random.default_rng(seed)
    B0 = rng.standard_normal((L, r))
    A0 = rng.standard_normal((M, r))
    return B0 @ A0.T + noise * rng.standard_normal((L, M))
