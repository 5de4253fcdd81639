"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): column shards + one packed NCCL all-reduce per
iteration must reproduce the oracle on the gathered problem (shard-count invariance)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_matches_oracle(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MGPU OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


# ------------------------------------------------------------------------------------------------ ONE process, several devices
def _single_process_cases():
    import numpy as np
    from oracle import vbmf_oracle as vo
    from tests.helpers import synth
    L, M, H = 96, 1001, 8
    Y = synth(L, M, 4, seed=5)
    return L, M, H, Y, vo, np


@pytest.mark.parametrize("ndev", [1, 2, 4, 8])
def test_single_process_multi_device(ndev):
    """vbmf_b200_mctx_*: the Julia drop-in is ONE process (src/vbmf.jl:175 is a plain call), so the column-sharded path must be
    reachable from it.  Full-size host arrays in, per-device host threads inside the library, results vs the unsharded oracle
    for all four parameter kinds (labels split by shard, Q2 map on global indices, trial row split across a shard boundary)."""
    if _ngpu() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    import copy
    from tests import gpu_helpers as G
    L, M, H, Y, vo, np = _single_process_cases()
    Yf = np.asfortranarray(Y)
    mc = G.vb.MultiContext(devices=list(range(ndev)))
    try:
        mc.attach(Yf)
        assert abs(mc.trYTY() - float(np.sum(Y * Y))) < 1e-10 * float(np.sum(Y * Y))
        assert sum(mc.shard(i)[1] for i in range(ndev)) == M
        niter = 6
        # dense with labels
        po = vo.vbmf_init(Y, H, H1=2, labels=list(range(3, M, 7)), rng=np.random.default_rng(11))
        q = G.to_gpu_params(po)
        _, it_o, d_o = vo.vbmf_run(Y, po, niter, eps=1e-12, est_covs=True, est_var=True)
        G.vb.vbmf_(None, q, niter, eps=1e-12, est_covs=True, est_var=True, ctx=mc)
        assert q.iterations == it_o
        G.compare(q, po, 1e-10)
        assert G.rel(q.YHat, po.BHat @ po.AHat.T) < 1e-10
        # sparse: diagonal path (Q2 needs global column indices) and full covariance + heteroscedastic + labels
        for ikw, kw in (({}, dict(full_cov=False, est_cb=True)), ({"H1": 2, "labels": list(range(3, M, 7))}, dict(full_cov=True, diag_var=True, est_cb=True))):
            po = vo.vbmf_sparse_init(Y, H, rng=np.random.default_rng(11), **ikw)
            q = G.to_gpu_params(po)
            d_o, it_o = vo.vbmf_sparse_run(Y, po, niter, eps=1e-12, **kw)
            d = G.vb.vbmf_sparse_(None, q, niter, eps=1e-12, ctx=mc, **kw)
            assert q.iterations == it_o and abs(d - d_o) <= 1e-8 * abs(d_o)
            G.compare(q, po, 1e-10)
            lb_o, lb = vo.sparse_lowerBound(Y, po), G.vb.lowerBound(None, q, ctx=mc)
            assert abs(lb - lb_o) <= 1e-10 * abs(lb_o)
        # dual with learned priors
        po = vo.vbmf_dual_init(Y, H, 3, rng=np.random.default_rng(11))
        q = G.to_gpu_params(po)
        d_o, it_o = vo.vbmf_dual_run(Y, po, niter, eps=1e-12, full_cov=False, est_priors=True, est_cb=True)
        d = G.vb.vbmf_dual_(None, q, niter, eps=1e-12, full_cov=False, est_priors=True, est_cb=True, ctx=mc)
        assert q.iterations == it_o
        G.compare(q, po, 1e-10)
        assert abs(G.vb.lowerBound(None, q, ctx=mc) - vo.dual_lowerBound(Y, po)) <= 1e-10 * abs(vo.dual_lowerBound(Y, po))
        # trial: the row split M0 falls inside a shard
        po = vo.vbmf_trial_init(Y, H, 3, 377, rng=np.random.default_rng(11))
        q = G.to_gpu_params(po)
        d_o, it_o = vo.vbmf_trial_run(Y, po, niter, eps=1e-12, full_cov=True, est_priors=True)
        d = G.vb.vbmf_trial_(None, q, niter, eps=1e-12, full_cov=True, est_priors=True, ctx=mc)
        assert q.iterations == it_o
        G.compare(q, po, 1e-10)
    finally:
        mc.close()


def test_two_devices_in_one_process_independent_contexts():
    """Two ordinary contexts on two different devices in one process (per-device kernel attributes): both must run the
    big-shared-memory kernels."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    from oracle import vbmf_oracle as vo
    from tests import gpu_helpers as G
    from tests.helpers import synth
    Y = synth(300, 2000, 16, seed=2)
    Yf = np.asfortranarray(Y)
    for dev in (1, 0, 1):
        c = G.vb.Context(device=dev)
        try:
            p = vo.vbmf_init(Y, 64, rng=np.random.default_rng(3))
            q = G.to_gpu_params(p)
            vo.vbmf_run(Y, p, 2, eps=0.0, est_covs=True, est_var=True)
            G.vb.vbmf_(Yf, q, 2, eps=0.0, est_covs=True, est_var=True, ctx=c, yhat=False)
            G.compare(q, p, 1e-10, ["AHat", "BHat", "SigmaA", "SigmaB", "sigma2"])
        finally:
            c.close()
