#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02n; mkdir -p $O
for f in 2 1 3; do
  VBMF_B200_K2_FOLD=$f timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=200 -k "contractions or wide_rank" > $O/pytest_fold$f.log 2>&1
  echo "fold=$f: $(tail -1 $O/pytest_fold$f.log)"
done
for f in 0 2 1 3; do
  VBMF_B200_K2_FOLD=$f timeout 400 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench_c5_fold$f.json 2> $O/bench_c5_fold$f.err
done
VBMF_B200_K2_FOLD=2 timeout 400 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout=300 -k "config5" > $O/pytest_fullsize_fold2.log 2>&1; tail -1 $O/pytest_fullsize_fold2.log
timeout 600 python tools/runs/probe_attach.py > $O/probe_attach.log 2>&1; cat $O/probe_attach.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02n/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("/")[-1], "value %.3f ms %.3f iterfrac %.4f k1 %.3f k2 %.3f"%(j["value"],j["ms_per_step"],r.get("iteration_frac_of_peak"),r["k1_ms"],r["k2_ms"]))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
