mkdir -p gpurun_out/r2j
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2j/tests.log 2>&1; tail -15 gpurun_out/r2j/tests.log
ncu --set full --clock-control none --import-source on -k regex:^chain_kernel -c 1 -o gpurun_out/r2j/cfg3_dmma_chain python scripts/cfg_once.py cfg3 1280 > gpurun_out/r2j/cfg3_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^chain_kernel -c 1 -o gpurun_out/r2j/cfg4_dmma_chain python scripts/cfg_once.py cfg4 1280 > gpurun_out/r2j/cfg4_full.log 2>&1
ls -la gpurun_out/r2j
