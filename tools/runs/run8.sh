#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02h; mkdir -p $O
K=tools/k4bench/k4bench
{ timeout 300 $K dmma 32 100000 20; timeout 300 $K reg 32 100000 20; } > $O/k4bench.jsonl 2> $O/k4bench.err
cat $O/k4bench.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
cp gpurun_out/parity_worst.json $O/ 2>/dev/null
grep -n "^FAILED\|passed\|failed" $O/pytest.log | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
for w in c4 c4full c5 c2; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02h/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("/")[-1], "value %.2f ms %.3f iterfrac %s e2e %s cpu %s"%(j["value"],j["ms_per_step"],r.get("iteration_frac_of_peak"), j["e2e"]["value"] if j.get("e2e") else None, j["cpu_baseline"]["value"] if j.get("cpu_baseline") else None))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-400:])
PY
