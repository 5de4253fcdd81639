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
    # the ctypes prototypes carry one argtype per C parameter
    plain = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for m in re.finditer(r"\b(?:int|int64_t|const char\*|void)\s+(vbmf_b200_\w+)\s*\(([^)]*)\)\s*;", plain):
        a = m.group(2).strip()
        assert len(vb._lib.SYMBOLS[m.group(1)][1]) == (0 if a in ("", "void") else len(a.split(","))), m.group(1)


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


def _c_struct_fields(header, name):
    import re
    m = re.search(r"typedef struct \{([^}]*)\} " + name + ";", header)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        mm = re.match(r"(const\s+)?(int64_t|double)\s*(\*?)\s*(.*)", decl)
        base, ptr, names = mm.group(2), mm.group(3), mm.group(4)
        for n in names.split(","):
            n = n.strip()
            p = ptr
            if n.startswith("*"):
                p, n = "*", n[1:].strip()
            out.append((n, ("Ptr" if p else "") + ("Int64" if base == "int64_t" else "Float64")))
    return out


def _julia_struct_fields(src, name):
    import re
    m = re.search(r"immutable " + name + r"\b[^\n]*\n(.*?)\nend", src, flags=re.S)
    out = []
    for line in m.group(1).split("\n"):
        for decl in line.split("#")[0].split(";"):
            decl = decl.strip()
            if decl:
                n, t = decl.split("::")
                out.append((n.strip(), t.strip().replace("Ptr{Int64}", "PtrInt64").replace("Ptr{Float64}", "PtrFloat64")))
    return out


def _call_arity(src, callee):
    """Number of top-level arguments of every `callee(` call in src."""
    out, i = [], 0
    while True:
        i = src.find(callee + "(", i)
        if i < 0:
            return out
        j, depth, n, seen = i + len(callee) + 1, 1, 0, False
        while depth:
            ch = src[j]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == "," and depth == 1:
                n += 1
            if not ch.isspace() and depth:
                seen = True
            j += 1
        out.append(n + 1 if seen else 0)
        i = j


def test_julia_shim_mirrors_the_header():
    """julia/VBMatrixFactorizationB200.jl cannot be executed here (no Julia): at least its four `immutable` struct mirrors must
    match include/vbmf_b200.h field for field (name, order, pointer-ness, integer/float), and every constructor call must pass
    exactly one value per field."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "vbmf_b200.h")).read()
    shim = open(os.path.join(root, "julia", "VBMatrixFactorizationB200.jl")).read()
    for cname, jname in (("vbmf_b200_dense_state", "DenseState"), ("vbmf_b200_sparse_state", "SparseState"),
                         ("vbmf_b200_dual_state", "DualState"), ("vbmf_b200_trial_state", "TrialState")):
        c, j = _c_struct_fields(header, cname), _julia_struct_fields(shim, jname)
        assert c == j, (cname, [(a, b) for a, b in zip(c, j) if a != b])
        body = shim.split("immutable " + jname, 1)[1].split("\nend", 1)[1]       # everything after the type definition
        calls = [n for n in _call_arity(body, jname) if n > 1]                      # skip Ref{T}-style mentions
        assert calls and all(n == len(c) for n in calls), (jname, calls, len(c))
    # every C symbol the shim ccalls is declared in the header
    import re
    for sym in set(re.findall(r"ccall\(\(:(\w+), LIB\)", shim)):
        assert re.search(r"\b" + sym + r"\s*\(", header), sym


def test_julia_shim_ccall_arities_match_prototypes():
    """Every `ccall((:sym, LIB), Cint, (types...), args...)` in the shim passes as many types and as many arguments as the C
    prototype of `sym` in include/vbmf_b200.h has parameters."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "vbmf_b200.h")).read(), flags=re.S)
    shim = open(os.path.join(root, "julia", "VBMatrixFactorizationB200.jl")).read()
    protos = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char\*|void)\s+(vbmf_b200_\w+)\s*\(([^)]*)\)\s*;", hdr):
        a = m.group(2).strip()
        protos[m.group(1)] = 0 if a in ("", "void") else len(a.split(","))

    def split_top(s):
        out, depth, cur = [], 0, ""
        for ch in s:
            depth += ch in "([{"
            depth -= ch in ")]}"
            if ch == "," and depth == 0:
                out.append(cur)
                cur = ""
            else:
                cur += ch
        return out + ([cur] if cur.strip() else [])
    i, n = 0, 0
    while True:
        i = shim.find("ccall((:", i)
        if i < 0:
            break
        j, depth = i + 6, 1
        while depth:
            depth += shim[j] in "([{"
            depth -= shim[j] in ")]}"
            j += 1
        parts = split_top(shim[i + 6:j - 1])
        sym = re.search(r":(\w+)", parts[0]).group(1)
        types = [t for t in split_top(parts[2].strip()[1:-1]) if t.strip()]
        assert protos.get(sym) == len(types) == len(parts) - 3, (sym, protos.get(sym), len(types), len(parts) - 3)
        i, n = j, n + 1
    assert n >= 15


def test_contraction_planner_host_logic(vb):
    """Split-K plans of K1 / K2 (host logic, no device needed): every slab is non-empty, the slabs cover L resp. M exactly once,
    K2 chunks are multiples of the 16-row stage, and at the BASELINE shapes the static work list fills >= 95 % of its last wave."""
    import ctypes as C
    lib = vb._lib.load()
    out = (C.c_int64 * 6)()

    def plan(L, M, H, sms=148):
        assert lib.vbmf_b200_plan_contractions(L, M, H, sms, C.cast(out, C.c_void_p)) == 0
        return list(out)
    rng = np.random.default_rng(0)
    shapes = [(20000, 200000, 64), (20000, 25000, 64), (10000, 100000, 32), (50000, 125000, 128), (10, 20, 2), (1, 1, 1), (38, 40, 20)]
    shapes += [(int(rng.integers(1, 60000)), int(rng.integers(1, 300000)), int(rng.integers(1, 129))) for _ in range(300)]
    for L, M, H in shapes:
        S1, kbs, S2, kchunk, cps, bn = plan(L, M, H)
        nkb = -(-L // 16)
        assert bn in (32, 64, 128) and bn >= H and (bn == 32 or bn // 2 < H)
        assert S1 >= 1 and kbs >= 1 and S1 * kbs >= nkb and (S1 - 1) * kbs < nkb, (L, M, H, S1, kbs)
        assert S2 >= 1 and kchunk % 16 == 0 and S2 * kchunk >= M and (S2 - 1) * kchunk < M, (L, M, H, S2, kchunk)
        assert S2 * H * ((L + 1) & ~1) * 8 <= (1 << 30) or S2 == 1          # K2 slab workspace cap
    for L, M, H in shapes[:4]:
        S1, kbs, S2, kchunk, cps, bn = plan(L, M, H)
        cap = 148 * cps
        for work in (-(-M // 128) * S1, -(-L // 128) * S2):
            assert work / (-(-work // cap) * cap) >= 0.95, (L, M, H, work, cap)
    assert lib.vbmf_b200_plan_contractions(10, 10, 129, 148, C.cast(out, C.c_void_p)) != 0      # H > 128 is refused


def test_logging_mirror_ignores_non_numeric_fields(vb):
    """A second logdir run on the same params must not trip over the trajectory stored by the first one (params.log) or over
    other non-numeric attributes; `failed` is not logged either."""
    p = vb.vbmf_init(np.zeros((4, 6)), 2, rng=np.random.default_rng(0))
    log = vb.create_log(p)
    p.log, p.failed, p.note = {"x": [1]}, False, "text"
    log2 = vb.create_log(p)
    assert set(log2) == set(log) and "log" not in log2 and "failed" not in log2 and "note" not in log2
    vb.update_log_(log2, p)
    assert all(len(v) == 2 for v in log2.values())


def test_a_copies_mismatch_detection(vb):
    """params.AHat = X without touching ATVecHat (examples/toy_data.jl:54): the step functions pick the copy their reference
    counterpart reads; lowerBound refuses the mixed state (needs no device: raised before any library call)."""
    from vbmf_b200 import api
    Y = np.random.default_rng(0).standard_normal((5, 7))
    p = vb.vbmf_sparse_init(Y, 3, rng=np.random.default_rng(1))
    assert not api._a_mismatch(p)
    p.AHat = np.asfortranarray(p.AHat + 1.0)
    assert api._a_mismatch(p)
    with pytest.raises(vb.VBMFError, match="disagree"):
        vb.lowerBound(Y, p)
    assert not api._a_mismatch(vb.vbmf_init(Y, 3))


def test_sharded_context_refuses_ambiguous_y(vb):
    """_ctx_for on a world > 1 context: a host Y of another shape than the attached shard cannot be placed."""
    from vbmf_b200 import api

    class FakeCtx:
        world, L, M, M_global, col_offset = 2, 5, 4, 8, 4
        attached = None

        def attach(self, Y, M_global=None, col_offset=0):
            self.attached = (Y.shape, M_global, col_offset)
    c = FakeCtx()
    with pytest.raises(vb.VBMFError, match="sharded context"):
        api._ctx_for(np.zeros((5, 8)), c)
    api._ctx_for(np.zeros((5, 4)), c)
    assert c.attached == ((5, 4), 8, 4)          # the shard keeps its global geometry
    assert api._ctx_for(None, c) is c


def test_peer_exchange_row_split_host_logic(vb):
    """Row split of the peer-exchange epilogue (host logic, no device needed): the ranks' 32-row tile ranges are contiguous,
    disjoint and cover every tile of BHat exactly once, no rank holds more than ceil(tiles / world), and the CTA count is the
    same on every rank and covers the largest share (or the cap of two CTAs per SM)."""
    import ctypes as C
    lib = vb._lib.load()
    out = (C.c_int64 * 3)()
    rng = np.random.default_rng(1)
    for L in [1, 31, 32, 33, 96, 517, 1000, 10000, 20000, 50000] + [int(x) for x in rng.integers(1, 200000, 60)]:
        ntiles = -(-L // 32)
        for world in range(1, 9):
            prev_hi, grids = 0, set()
            for rank in range(world):
                assert lib.vbmf_b200_px_plan(L, world, rank, C.cast(out, C.c_void_p)) == 0
                lo, hi, grid = list(out)
                assert lo == prev_hi and lo <= hi <= ntiles and hi - lo <= -(-ntiles // world), (L, world, rank, lo, hi)
                prev_hi = hi
                grids.add(grid)
                assert grid >= 1 and (grid >= hi - lo or grid == 296)
            assert prev_hi == ntiles and len(grids) == 1, (L, world, prev_hi, grids)
    assert lib.vbmf_b200_px_plan(100, 9, 0, C.cast(out, C.c_void_p)) != 0       # more than 8 ranks use the NCCL path
    assert lib.vbmf_b200_px_plan(100, 2, 2, C.cast(out, C.c_void_p)) != 0


def test_log_helpers_survive_a_second_logged_run(vb, tmp_path):
    """create_log / update_log_ / save_log / load_log are host logic (src/data_manip.jl:6-118).  A logged run leaves the
    trajectory on the params object (`params.log`); creating a log from the SAME object again -- resuming with logging -- must
    skip it (round 1 crashed there with a TypeError), and only numeric fields are logged."""
    rng = np.random.default_rng(0)
    Y = rng.standard_normal((6, 9))
    for p in (vb.vbmf_init(Y, 3, rng=rng), vb.vbmf_sparse_init(Y, 3, rng=rng), vb.vbmf_dual_init(Y, 3, 1, rng=rng)):
        log = vb.create_log(p)
        assert "log" not in log and "YHat" not in log and "BHat" in log
        p.BHat = p.BHat + 1.0
        vb.update_log_(log, p)
        p.log = log                       # what a logged run leaves behind
        p.failed = False
        p.desc = "text is not logged"
        log2 = vb.create_log(p)           # second logged run on the same params
        assert set(log2) == set(log)
        vb.update_log_(log2, p)
        d = vb.save_log(log2, Y, {"ca": 1.0}, str(tmp_path), desc="twice_%s" % p.kind)
        back, Yback = vb.load_log(d)
        assert back["BHat"].shape == p.BHat.shape + (2,) and np.array_equal(back["BHat"][..., 1], p.BHat)
        assert np.array_equal(Yback, Y)
