# one GPU: the exchange kernels against themselves (W = 1) on one 25000-column shard, with the segment / phase profile
VBMF_B200_PX_SELF=1 timeout 300 python bench.py --workload c3shard8 --no-e2e --no-cpu --segments 2>/dev/null | tail -1 > gpurun_out/self_segments.json
python - <<'PY'
import json
j=json.loads(open("gpurun_out/self_segments.json").read().strip().splitlines()[-1]); r=j["roofline"]
print(round(j["ms_per_step"],4), {k: round(v*1e3,1) for k,v in r.get("segments_ms").items() if not k.startswith("k")})
PY
