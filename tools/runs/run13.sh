#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02m; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
grep -n "^FAILED\|passed\|failed" $O/pytest.log | tail -12
cp gpurun_out/parity_worst.json $O/ 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > $O/bench_c3_overlap.json 2> $O/bench_c3_overlap.err
VBMF_B200_NO_UPLOAD_OVERLAP=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > $O/bench_c3_no_overlap.json 2> $O/bench_c3_no_overlap.err
VBMF_B200_ATTACH_CHUNK_MB=0 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > $O/bench_c3_sync_attach.json 2> $O/bench_c3_sync_attach.err
timeout 900 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu > $O/bench_c5.json 2> $O/bench_c5.err
timeout 900 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu > $O/bench_c4.json 2> $O/bench_c4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02m/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j["roofline"]
        print(f.split("/")[-1], "value %.2f ms %.3f iterfrac %.4f"%(j["value"],j["ms_per_step"],r.get("iteration_frac_of_peak")), "e2e", j["e2e"]["value"], "s", j["e2e"]["seconds"])
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
