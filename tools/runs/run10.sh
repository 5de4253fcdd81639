#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02j; mkdir -p $O
K=tools/k4bench/k4bench
for H in 12 20 32; do timeout 60 $K hxh $H 50; VBMF_B200_HXH=warp timeout 60 $K hxh $H 50; done > $O/hxh.jsonl 2> $O/hxh.err
cat $O/hxh.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
grep -n "^FAILED\|passed\|failed" $O/pytest.log | tail -12
for w in c3shard8 c3 c4 c4full; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-e2e > $O/bench_${w}_sk.json 2> $O/bench_${w}_sk.err
  VBMF_B200_K2=classic timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-e2e > $O/bench_${w}_classic.json 2> $O/bench_${w}_classic.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02j/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("/")[-1], "value %.2f ms %.3f iterfrac %.4f k1 %.3f k2 %.3f"%(j["value"],j["ms_per_step"],r.get("iteration_frac_of_peak"), r["k1_ms"], r["k2_ms"]))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-400:])
PY
