#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02c; mkdir -p $O
K=tools/k4bench/k4bench
{
  timeout 120 $K dmma 32 4096 5
  timeout 120 $K dmma 24 4096 5
  timeout 120 $K dmma 16 4096 5
  timeout 120 $K dmma 7 4096 5
  VBMF_B200_K4_MINB=4 timeout 120 $K dmma 32 4096 5
  VBMF_B200_K4_MINB=5 timeout 120 $K dmma 32 4096 5
  timeout 300 $K dmma 32 100000 20
  VBMF_B200_K4_MINB=4 timeout 300 $K dmma 32 100000 20
  VBMF_B200_K4_MINB=5 timeout 300 $K dmma 32 100000 20
  VBMF_B200_K4_MINB=2 timeout 300 $K dmma 32 100000 20
  timeout 300 $K dmma 16 100000 20
} > $O/k4bench.jsonl 2> $O/k4bench.err
for H in 5 20 32 50 64 100 128; do timeout 60 $K hxh $H 50; done > $O/hxh.jsonl 2>> $O/k4bench.err
cat $O/k4bench.jsonl $O/hxh.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sparse_A_full -s 2 -c 1 -f -o $O/k4_dmma3 $K dmma 32 100000 1 > $O/ncu3.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
cp gpurun_out/parity_worst.json $O/ 2>/dev/null
grep -n "^FAILED\|passed\|failed" $O/pytest.log | tail -12
timeout 600 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu > $O/bench_c4.json 2> $O/bench_c4.err
timeout 600 python bench.py --workload c4full --steps 20 --warmup 3 --no-cpu > $O/bench_c4full.json 2> $O/bench_c4full.err
python - <<'PY'
import json
for w in ("c4","c4full"):
    try:
        j=json.loads(open("gpurun_out/r02c/bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, "ms/step %.3f"%j["ms_per_step"], "iter frac %.3f"%j["roofline"]["iteration_frac_of_peak"], "e2e", j["e2e"]["value"] if j.get("e2e") else None)
    except Exception as e: print(w, "ERR", e)
PY
