for nc in 0; do
if [ $nc = 1 ]; then export VBMF_B200_NO_CARVEOUT=1; fi
VBMF_B200_PX_STAMPS=1 VBMF_B200_PX_SELF=1 timeout 300 python bench.py --workload c3shard8 --no-e2e --no-cpu --segments 2>gpurun_out/self_$nc.err | tail -1 > gpurun_out/self_$nc.json
grep "px " gpurun_out/self_$nc.err | tail -3
python - <<PY
import json
j=json.loads(open("gpurun_out/self_$nc.json").read().strip().splitlines()[-1]); r=j["roofline"]
print("no_carveout=$nc", round(j["ms_per_step"],4), {k: round(v*1e3,1) for k,v in r.get("segments_ms").items() if not k.startswith("k")})
PY
timeout 300 python bench.py --workload c3shard8 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('plain no_carveout=$nc', round(j['ms_per_step'],4))"
done
