#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02i; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
grep -n "^FAILED\|passed\|failed" $O/pytest.log | tail -12
for w in c3shard8 c3 c4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-e2e > $O/bench_${w}_sk.json 2> $O/bench_${w}_sk.err
  VBMF_B200_K2=classic timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-e2e > $O/bench_${w}_classic.json 2> $O/bench_${w}_classic.err
done
VBMF_B200_K2=streamk timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench_c5_sk.json 2> $O/bench_c5_sk.err
timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench_c5_classic.json 2> $O/bench_c5_classic.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3shard8.csv python bench.py --workload c3shard8 --steps 3 --warmup 3 --no-cpu --no-e2e > $O/ncu_c3shard8.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02i/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("/")[-1], "value %.2f ms %.3f iterfrac %.4f k1 %.3f k2 %.3f"%(j["value"],j["ms_per_step"],r.get("iteration_frac_of_peak"), r["k1_ms"], r["k2_ms"]))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-400:])
PY
python tools/launch_summary.py $O/launches_c3shard8.csv
