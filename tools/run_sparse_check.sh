# one GPU: sparse / dual / trial parity, then the three non-dense workloads
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_peer_exchange.py -x -q -k "sparse or dual or trial or peer or fuzz or lower or wide" 2>&1 | tail -3
for w in c4 c4full c5; do timeout 200 python bench.py --workload $w --no-cpu --no-e2e 2>/dev/null | tail -1 > gpurun_out/s6_$w.json; done
python - <<'PY'
import json
for w in ("c4","c4full","c5"):
    j=json.loads(open("gpurun_out/s6_%s.json"%w).read().strip().splitlines()[-1]); r=j["roofline"]
    print(w, round(j["ms_per_step"],4), round(r["iteration_frac_of_peak"],4), j["clocks"]["sm_mhz"], j["clocks"]["samples"])
PY
