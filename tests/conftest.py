import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _parity_recorder(request):
    """Names the running test for tests/gpu_helpers.compare's worst-error record."""
    mod = sys.modules.get("tests.gpu_helpers")
    if mod is not None:
        mod.CURRENT[0] = request.node.nodeid
    yield
    mod = sys.modules.get("tests.gpu_helpers")
    if mod is not None:
        mod.CURRENT[0] = request.node.nodeid


def pytest_sessionfinish(session, exitstatus):
    """GPU runs: the measured worst relative error of every parity comparison, per test and field (copied to profiles/)."""
    mod = sys.modules.get("tests.gpu_helpers")
    if mod is None or not mod.WORST:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        worst = {k: v for k, v in mod.WORST.items() if v}
        summary = {"tests": len(worst), "max_rel_err": max((max(v.values()) for v in worst.values()), default=0.0)}
        with open(os.path.join(out, "parity_worst.json"), "w") as f:
            json.dump({"summary": summary, "per_test": worst}, f, indent=1, sort_keys=True)
    except OSError:
        pass
