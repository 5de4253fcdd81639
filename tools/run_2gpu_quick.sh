TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_peer_exchange.py -x -q 2>&1 | tail -3
timeout 300 $TR --master-port 29611 tests/mgpu_worker.py 2>&1 | grep -E "MGPU|Error|error|assert" | head -5
bash tools/run_2gpu_phases.sh
