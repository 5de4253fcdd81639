# 2-GPU validation of the peer exchange: parity (one process per GPU and single-process multi-device), then the A/B of the
# exchange against the NCCL all-reduce at the 8-GPU shard size (25000 columns per GPU) with segment timings.
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_peer_exchange.py -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15
timeout 300 $TR --master-port 29611 tests/mgpu_worker.py 2>&1 | grep -E "MGPU|Error|error|assert" | head -5 | tee gpurun_out/t2_mgpu_worker_n2.log
bash tools/run_2gpu_ab.sh 2>&1 | grep "^gpurun_out"
