#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02g; mkdir -p $O
K=tools/k4bench/k4bench
{ timeout 120 $K dmma 32 4096 5; timeout 300 $K dmma 32 100000 20; } > $O/k4bench.jsonl 2> $O/k4bench.err
cat $O/k4bench.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
cp gpurun_out/parity_worst.json $O/ 2>/dev/null
grep -n "^FAILED\|passed\|failed" $O/pytest.log | tail -12
for w in c4 c4full c5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench_$w.json 2> $O/bench_$w.err
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e > $O/ncu_$w.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02g/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("/")[-1], "it/s %.2f ms %.3f iterfrac %.3f k1 %.3f k2 %.3f share %.3f"%(j["value"],j["ms_per_step"],r["iteration_frac_of_peak"],r["k1_ms"],r["k2_ms"],r["contraction_share_of_step"]))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-400:])
PY
for w in c4 c4full c5; do echo "== $w"; python tools/launch_summary.py $O/launches_$w.csv; done
