"""CPU-side checks of the boundary: the shared library builds, loads and exports every symbol include/vbmf_b200.h declares;
the product path refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import vbmf_b200_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vb():
    vbmf_b200_loader.build()
    return vbmf_b200_loader.load()


def test_header_symbols_exported(vb):
    hdr = open(os.path.join(ROOT, "include", "vbmf_b200.h")).read()
    declared = set(re.findall(r"\b(vbmf_b200_[a-zA-Z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = ctypes.CDLL(vb.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(vb._lib.SYMBOLS), declared ^ set(vb._lib.SYMBOLS)


def test_struct_sizes_match_header(vb):
    # 64-bit fields only, so ctypes and the C compiler agree on the layout; guard against drift in the field count
    hdr = open(os.path.join(ROOT, "include", "vbmf_b200.h")).read()
    for cname, cls in (("vbmf_b200_dense_state", vb._lib.DenseState), ("vbmf_b200_sparse_state", vb._lib.SparseState),
                       ("vbmf_b200_dual_state", vb._lib.DualState), ("vbmf_b200_trial_state", vb._lib.TrialState)):
        body = re.search(r"typedef struct \{([^}]*)\} %s;" % cname, hdr).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            decl = re.sub(r"^(const\s+)?(int64_t|double)\s*\*?", "", stmt)
            names += [n.strip().lstrip("*").strip() for n in decl.split(",")]
        assert names == [f[0] for f in cls._fields_], (cname, names)
        assert ctypes.sizeof(cls) == 8 * len(names)


def test_no_cpu_fallback(vb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert vb._lib.load().vbmf_b200_device_count() == 0
    with pytest.raises(vb.VBMFError, match="no CPU fallback"):
        vb.Context()
    with pytest.raises(vb.VBMFError):
        vb.vbmf_(np.zeros((4, 5), order="F"), vb.vbmf_init(np.zeros((4, 5)), 2), 1)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "vbmatrixfactorization.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)


def test_init_mirrors_reference(vb):
    Y = np.arange(12.0).reshape(3, 4)
    p = vb.vbmf_init(Y, 2, ca=0.5, H1=1, labels=[2, 4], rng=np.random.default_rng(0))
    assert p.AHat.shape == (4, 2) and p.BHat.shape == (3, 2) and np.all(p.AHat[[1, 3], 1] == 0.0)
    assert np.allclose(p.invCA, 2 * np.eye(2)) and p.sigma2 == 1.0
    s = vb.vbmf_sparse_init(Y, 2, rng=np.random.default_rng(0))
    assert s.alpha == 1e-10 + 0.5 and s.gamma == 1e-10 + 1.5 and s.eta == 1e-10 + 6.0 and s.trYTY == float((Y * Y).sum())
    assert np.array_equal(s.ATVecHat, np.ascontiguousarray(s.AHat).reshape(-1))
    d = vb.vbmf_dual_init(Y, 3, 1, rng=np.random.default_rng(0))
    assert d.A0Hat.shape == (4, 1) and d.A1Hat.shape == (4, 2) and d.CA0.shape == (4,) and d.CA1.shape == (8,)
    with pytest.raises(vb.VBMFError, match="H must be at least H0"):
        vb.vbmf_dual_init(Y, 1, 2)
    assert vb.shard_columns(10, 4, 0) == (0, 3) and vb.shard_columns(10, 4, 3) == (8, 2)
