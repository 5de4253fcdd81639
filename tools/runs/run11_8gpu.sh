#!/bin/bash
# 8-GPU box: sharded parity at 4 / 8 ranks, single-process multi-device at 4 / 8 devices, scaling bench lines with state_check
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02k; mkdir -p $O
nvidia-smi -L > $O/smi.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29711 tests/mgpu_worker.py > $O/mgpu_worker_n8.log 2>&1; grep "MGPU" $O/mgpu_worker_n8.log
timeout 300 $TR --nproc-per-node 4 --master-port 29712 tests/mgpu_worker.py > $O/mgpu_worker_n4.log 2>&1; grep "MGPU" $O/mgpu_worker_n4.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=600 -k "single_process and (8 or 4) or two_devices" > $O/pytest_multi.log 2>&1; tail -3 $O/pytest_multi.log
timeout 600 $TR --nproc-per-node 8 --master-port 29713 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > $O/bench_c3_n8.json 2> $O/bench_c3_n8.err
timeout 600 $TR --nproc-per-node 4 --master-port 29714 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu --no-e2e > $O/bench_c3_n4.json 2> $O/bench_c3_n4.err
timeout 600 python bench.py --gpus 8 --single-process --steps 20 --warmup 3 > $O/bench_c3_n8_single_process.json 2> $O/bench_c3_n8_single_process.err
timeout 900 $TR --nproc-per-node 8 --master-port 29715 bench.py --gpus 8 --workload c5 --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench_c5_n8.json 2> $O/bench_c5_n8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02k/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j.get("roofline",{})
        print(f.split("/")[-1], "it/s %.2f"%j["value"], "ms %.3f"%j["ms_per_step"], "share", r.get("contraction_share_of_step"), "e2e", (j.get("e2e") or {}).get("value"), j.get("state_check"))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
