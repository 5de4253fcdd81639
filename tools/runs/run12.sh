#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02l; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
grep -n "^FAILED\|passed\|failed" $O/pytest.log | tail -12
cp gpurun_out/parity_worst.json $O/ 2>/dev/null
for w in c4 c4full; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-e2e > $O/bench_$w.json 2> $O/bench_$w.err
done
VBMF_B200_NO_DIAG_FUSION=1 timeout 600 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu --no-e2e > $O/bench_c4_nofusion.json 2> $O/bench_c4_nofusion.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"sparse_A_diag_fused|sparse_diag_reduce|y_stats|sum_partials|B_epilogue|gram_dmma|hxh|B_reduce|update_CA" -c 40 --csv --log-file $O/hbm_kernels_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_hbm_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_y -s 6 -c 2 -f -o $O/gemm_c3 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_gemm_c3.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02l/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("/")[-1], "value %.2f ms %.3f iterfrac %.4f k1 %.3f k2 %.3f"%(j["value"],j["ms_per_step"],r.get("iteration_frac_of_peak"), r["k1_ms"], r["k2_ms"]))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-400:])
PY
