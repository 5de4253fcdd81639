# final one-GPU validation: GPU suite, smoke, the headline bench (both arms), the other workloads, launch lists
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/fin_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 300 python bench.py 2>/dev/null | tail -1 > gpurun_out/fin_c3.json; cut -c1-200 gpurun_out/fin_c3.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/fin_ref.json; cut -c1-200 gpurun_out/fin_ref.json
for w in c3shard8 c4 c4full c5 c2; do timeout 300 python bench.py --workload $w 2>/dev/null | tail -1 > gpurun_out/fin_$w.json; cut -c1-200 gpurun_out/fin_$w.json; done
for w in c3 c3shard8; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fin_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/fin_launches_$w.csv | tee gpurun_out/fin_launches_${w}_summary.txt
done
