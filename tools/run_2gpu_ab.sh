# A/B of the peer exchange against the NCCL all-reduce at the 8-GPU shard size (25000 columns per GPU), with segment timings
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29612 bench.py --gpus 2 --workload c3quarter --no-e2e --no-cpu --segments 2>gpurun_out/t2_q_px.err | tail -1 > gpurun_out/t2_c3quarter_n2_px.json
VBMF_B200_NO_PX=1 timeout 300 $TR --master-port 29613 bench.py --gpus 2 --workload c3quarter --no-e2e --no-cpu --segments 2>gpurun_out/t2_q_nccl.err | tail -1 > gpurun_out/t2_c3quarter_n2_nccl.json
VBMF_B200_PX_SELF=1 timeout 300 python bench.py --workload c3shard8 --no-e2e --no-cpu --segments 2>/dev/null | tail -1 > gpurun_out/t2_c3shard8_self.json
timeout 300 python bench.py --workload c3shard8 --no-e2e --no-cpu --segments 2>/dev/null | tail -1 > gpurun_out/t2_c3shard8.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/t2_c3*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j["roofline"]
        print(f, round(j["ms_per_step"],4), r.get("peer_exchange"), r.get("segments_ms"))
    except Exception as e: print(f, "ERR", e)
PY
# headline mode (no segment marks): what the driver's run sees
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29615 bench.py --gpus 2 --workload c3quarter --no-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/t2_plain_c3quarter_n2_px.json
python -c "
import json
j=json.loads(open('gpurun_out/t2_plain_c3quarter_n2_px.json').read().strip().splitlines()[-1]); print('plain px', round(j['ms_per_step'],4), j['roofline'].get('allreduce_ms'))"
