#!/bin/bash
# GPU round script 1: K4 microbenchmark variants, full GPU test suite, short bench lines.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/smi.txt 2>&1
K=tools/k4bench/k4bench
{
  timeout 120 $K dmma 32 4096 5
  timeout 120 $K dmma 24 4096 5
  timeout 120 $K dmma 16 4096 5
  timeout 120 $K dmma 8 4096 5
  timeout 120 $K dmma 5 4096 5
  timeout 120 $K reg 32 4096 5
  timeout 300 $K dmma 32 100000 20
  VBMF_B200_K4_MINB=2 timeout 300 $K dmma 32 100000 20
  timeout 300 $K reg 32 100000 20
  timeout 300 $K dmma 16 100000 20
  timeout 300 $K reg 16 100000 20
} > $O/k4bench.jsonl 2> $O/k4bench.err
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
cp gpurun_out/parity_worst.json $O/ 2>/dev/null
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_c3.json 2> $O/bench_c3.err
timeout 600 python bench.py --workload c4full --steps 20 --warmup 3 --no-cpu > $O/bench_c4full.json 2> $O/bench_c4full.err
timeout 600 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu > $O/bench_c4.json 2> $O/bench_c4.err
tail -3 $O/pytest.log; cat $O/k4bench.jsonl; tail -c 600 $O/bench_c4full.json
