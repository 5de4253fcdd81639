# diagnosis: phases of the exchange epilogue / Gram reduction at W = 2 (25000 columns per GPU) and W = 1 (self)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29612 bench.py --gpus 2 --workload c3quarter --no-e2e --no-cpu --segments 2>gpurun_out/t2_q_px.err | tail -1 > gpurun_out/t2_c3quarter_n2_px.json
VBMF_B200_PX_SELF=1 timeout 300 python bench.py --workload c3shard8 --no-e2e --no-cpu --segments 2>/dev/null | tail -1 > gpurun_out/t2_c3shard8_self.json
python - <<'PY'
import json
for f in ("gpurun_out/t2_c3quarter_n2_px.json", "gpurun_out/t2_c3shard8_self.json"):
    j=json.loads(open(f).read().strip().splitlines()[-1]); r=j["roofline"]
    print(f, round(j["ms_per_step"],4), {k: round(v*1e3,1) for k,v in r.get("segments_ms").items() if not k.startswith("k")})
PY
