"""CPU checks of the bench.py contract that need no GPU: the reference arm prints one JSON line with the agreed keys, and the
logging mirror round-trips."""
import json
import os
import subprocess
import sys

import numpy as np

import vbmf_b200_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, VBMF_BENCH_CPU_BUDGET_S="3")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "iterations/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "20000x200000x64" in d["metric"] and "workload" in d["config"]


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_log_roundtrip(tmp_path):
    vb = vbmf_b200_loader.load()
    Y = np.arange(12.0).reshape(3, 4)
    p = vb.vbmf_init(Y, 2, rng=np.random.default_rng(0))
    log = vb.create_log(p)
    p.sigma2 = 0.5
    p.BHat = p.BHat + 1.0
    vb.update_log_(log, p)
    d = vb.save_log(log, Y, {}, str(tmp_path), desc="t")
    lg, Yl = vb.load_log(d)
    assert np.array_equal(Yl, Y) and lg["sigma2"].tolist() == [1.0, 0.5] and lg["BHat"].shape == (3, 2, 2)
    q = vb.vbmf_init(Y, 2, rng=np.random.default_rng(1))
    vb.extract_params_(lg, q, 1)
    assert q.sigma2 == 0.5 and np.array_equal(q.BHat, p.BHat)
