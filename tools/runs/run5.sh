#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02e; mkdir -p $O
for w in c3shard8 c4 c4full; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-e2e > $O/bench_$w.json 2> $O/bench_$w.err
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e > $O/ncu_$w.log 2>&1
done
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 900 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu > $O/bench_c5.json 2> $O/bench_c5.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02e/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("/")[-1], "it/s %.2f ms %.3f iterfrac %.3f k1 %.3f k2 %.3f share %.3f e2e %s"%(j["value"],j["ms_per_step"],r["iteration_frac_of_peak"],r["k1_ms"],r["k2_ms"],r["contraction_share_of_step"], j["e2e"]["value"] if j.get("e2e") else None))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-400:])
PY
for w in c3shard8 c4 c4full; do echo "== $w"; python tools/launch_summary.py $O/launches_$w.csv; done
