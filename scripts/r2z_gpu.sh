# final-state evidence of round 2: tests, bench lines of every configuration, launch list, ncu captures of the chain kernel and the sampler
mkdir -p gpurun_out/r2z
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2z/gpu_tests.log 2>&1; tail -3 gpurun_out/r2z/gpu_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2z/bench_cfg2.json 2> gpurun_out/r2z/bench_cfg2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z/bench_cfg2_reference_arm.json 2> gpurun_out/r2z/bench_ref.err
for c in cfg1 cfg3 cfg4 cfg5 cfg2p; do
  python bench.py --config $c --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/r2z/bench_$c.json 2> gpurun_out/r2z/bench_$c.err || tail -5 gpurun_out/r2z/bench_$c.err
done
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2z/b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z/cfg2_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2z/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 2 -c 1 -o gpurun_out/r2z/cfg2_tc16p_chain python scripts/tc_once.py 128 > gpurun_out/r2z/cfg2_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 0 -c 1 -o gpurun_out/r2z/cfg2_tc_sampler python scripts/tc_once.py 1280 > gpurun_out/r2z/cfg2_sampler.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:chain_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2z/chain_traffic_ns10000.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2z/traffic.log 2>&1
python scripts/base_determinism.py 10000 3 > gpurun_out/r2z/determinism.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2z/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if d.get('impl') == 'reference':
            print(f.split('/')[-1], d['value'], d['unit'], d.get('cpu_baseline')); continue
        r=d['roofline']
        print(f.split('/')[-1], 'value %.1f ms/step %.2f e2e %.1f | chain %.2f ms %.2f TF frac %.3f (%s peak %.1f) | cpu %s | stages %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel_ms'], r['achieved'], r['frac'], r['bound'], r['peak'], (d.get('cpu_baseline') or {}).get('value'), {k:round(v,2) for k,v in d['stages_ms'].items()}))
    except Exception as e:
        print(f, 'FAILED', e)
PY
ls -la gpurun_out/r2z
