# one-GPU check: full GPU suite, then every workload with the per-segment timings of one iteration
set -x
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6)
timeout 300 python bench.py --no-cpu --no-e2e 2>gpurun_out/s2_c3.err | tail -1 > gpurun_out/s2_c3.json
for w in c3shard8 c4 c4full c5; do timeout 300 python bench.py --workload $w --no-cpu --no-e2e 2>/dev/null | tail -1 > gpurun_out/s2_$w.json; done
VBMF_B200_PX_SELF=1 timeout 300 python bench.py --workload c3shard8 --no-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/s2_c3shard8_self.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s2_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j["roofline"]
        print(f, round(j["ms_per_step"],4), r.get("iteration_frac_of_peak"), r.get("segments_ms"))
    except Exception as e: print(f, "ERR", e)
PY
