#!/bin/bash
# 2-GPU box: multi-GPU parity (torchrun workers + single-process multi-device ABI), sharded bench lines with state_check
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02d; mkdir -p $O
nvidia-smi -L > $O/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=600 > $O/pytest_multi.log 2>&1
echo "rc=$?" >> $O/pytest_multi.log; tail -5 $O/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_worker.py > $O/mgpu_worker_n2.log 2>&1
grep "MGPU" $O/mgpu_worker_n2.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench_c3_n1.json 2> $O/bench_c3_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench_c3_n2.json 2> $O/bench_c3_n2.err
timeout 600 python bench.py --gpus 2 --single-process --steps 10 --warmup 3 > $O/bench_c3_n2_single_process.json 2> $O/bench_c3_n2_single_process.err
timeout 600 python bench.py --gpus 1 --single-process --steps 10 --warmup 3 > $O/bench_c3_n1_single_process.json 2> $O/bench_c3_n1_single_process.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02d/bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "it/s %.2f"%j["value"], "ms %.3f"%j["ms_per_step"], j.get("state_check"))
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-400:])
PY
