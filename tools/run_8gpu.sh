# 8-GPU validation of the peer exchange (parity at W = 8 in both process models) and the strong-scaling A/B against NCCL
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29711 tests/mgpu_worker.py 2>&1 | grep -E "MGPU|Error|error|assert" | head -5 | tee gpurun_out/t8_mgpu_worker_n8.log
timeout 200 $TR --nproc-per-node 8 --master-port 29712 bench.py --gpus 8 --no-cpu 2>gpurun_out/t8_px.err | tail -1 > gpurun_out/t8_c3_n8_px.json
timeout 200 $TR --nproc-per-node 8 --master-port 29715 bench.py --gpus 8 --no-e2e --no-cpu --segments 2>/dev/null | tail -1 > gpurun_out/t8_c3_n8_px_segments.json
VBMF_B200_NO_PX=1 timeout 200 $TR --nproc-per-node 8 --master-port 29713 bench.py --gpus 8 --no-e2e --no-cpu 2>gpurun_out/t8_nccl.err | tail -1 > gpurun_out/t8_c3_n8_nccl.json
timeout 200 $TR --nproc-per-node 4 --master-port 29714 bench.py --gpus 4 --no-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/t8_c3_n4_px.json
timeout 200 $TR --nproc-per-node 2 --master-port 29716 bench.py --gpus 2 --no-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/t8_c3_n2_px.json
timeout 200 python bench.py --gpus 1 --no-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/t8_c3_n1.json
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "single_process and 8" 2>&1 | tail -3 | tee gpurun_out/t8_pytest_single_process_8.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/t8_c3*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j.get("roofline") or {}
        seg={k: round(v*1e3,1) for k,v in (r.get("segments_ms") or {}).items() if not k.startswith("k")}
        print(f, j.get("n_gpus"), round(j["ms_per_step"],4), round(j["value"],2), r.get("peer_exchange"), seg, (j.get("e2e") or {}).get("value"), j.get("state_check",{}).get("norm_BHat_fro"))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/t8_px.err
