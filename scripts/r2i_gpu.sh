mkdir -p gpurun_out/r2i
ncu --set full --clock-control none --import-source on -k regex:f64mma -c 1 -o gpurun_out/r2i/cfg3_dmma_chain python scripts/cfg_once.py cfg3 1280 > gpurun_out/r2i/cfg3_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mdmma -c 1 -o gpurun_out/r2i/cfg4_dmma_chain python scripts/cfg_once.py cfg4 1280 > gpurun_out/r2i/cfg4_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 1 -c 1 -o gpurun_out/r2i/cfg2_tc16p_chain python scripts/tc_once.py 128 > gpurun_out/r2i/cfg2_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gru_sample_kernel -c 1 -o gpurun_out/r2i/cfg2_sampler python scripts/tc_once.py 1280 > gpurun_out/r2i/cfg2_sampler.log 2>&1
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2i/b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i/cfg2_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2i/ncu_b.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:chain_kernel -s 1 -c 1 --csv --log-file gpurun_out/r2i/chain_traffic_ns10000.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2i/traffic.log 2>&1
for c in cfg3 cfg4; do python bench.py --config $c --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/r2i/bench_$c.json 2> gpurun_out/r2i/bench_$c.err; done
ls -la gpurun_out/r2i
