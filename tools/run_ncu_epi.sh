# ncu source-level capture of the exchange epilogue (single rank, self exchange) -- stall reasons per instruction
export VBMF_B200_PX_SELF=1
timeout 300 python bench.py --workload c3shard8 --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1 || exit 1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:B_epilogue_dmma_kernel -s 3 -c 1 -o gpurun_out/epi_px -f python bench.py --workload c3shard8 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_epi.log 2>&1
ls -la gpurun_out/epi_px.ncu-rep
