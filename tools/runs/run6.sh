#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02f; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
cp gpurun_out/parity_worst.json $O/ 2>/dev/null
grep -n "^FAILED\|passed\|failed" $O/pytest.log | tail -12
for w in c4 c4full; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu > $O/bench_$w.json 2> $O/bench_$w.err
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e > $O/ncu_$w.log 2>&1
done
# HBM-bound kernels: DRAM bytes and duration (one capture each)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"sparse_A_diag_fused|update_CA|sparse_diag_reduce|y_stats|sum_partials|dense_A_fused|B_epilogue|gram_dmma" -c 60 --csv --log-file $O/hbm_kernels_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_hbm_c4.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"sparse_A_diag_fused|update_CA|sparse_diag_reduce|y_stats|sum_partials|B_epilogue|gram_dmma|lb_partial" -c 60 --csv --log-file $O/hbm_kernels_c5.csv python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_hbm_c5.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02f/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("/")[-1], "it/s %.2f ms %.3f iterfrac %.3f k1 %.3f k2 %.3f share %.3f e2e %s clocks %s"%(j["value"],j["ms_per_step"],r["iteration_frac_of_peak"],r["k1_ms"],r["k2_ms"],r["contraction_share_of_step"], j["e2e"]["value"] if j.get("e2e") else None, j["clocks"]))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-400:])
PY
for w in c4 c4full; do echo "== $w"; python tools/launch_summary.py $O/launches_$w.csv; done
