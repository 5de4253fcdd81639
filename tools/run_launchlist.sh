# per-kernel durations of one iteration (ncu launch list; numbers under ncu are never bench values)
set -x
W=${1:-c3shard8}
timeout 300 python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$W.csv python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_$W.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$W.csv
