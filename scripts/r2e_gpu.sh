mkdir -p gpurun_out/r2e
for c in cfg1 cfg3 cfg4 cfg5 cfg2p; do
  python bench.py --config $c --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/r2e/bench_$c.json 2> gpurun_out/r2e/bench_$c.err || tail -5 gpurun_out/r2e/bench_$c.err
done
python bench.py --steps 5 --warmup 3 > gpurun_out/r2e/bench_cfg2.json 2> gpurun_out/r2e/bench_cfg2.err
ncu --set full --clock-control none --import-source on -k regex:md_chain_kernel -c 1 -o gpurun_out/r2e/cfg4_chain python scripts/cfg_once.py cfg4 1280 > gpurun_out/r2e/cfg4_full.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2e/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f.split('/')[-1], 'value %.1f ms/step %.2f e2e %.1f | chain %.2f ms %.2f TF frac %.3f (%s peak %.1f) | cpu %.3f | stages %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel_ms'], r['achieved'], r['frac'], r['bound'], r['peak'], d['cpu_baseline']['value'], {k:round(v,2) for k,v in d['stages_ms'].items()}))
        if 'fp64_dfma_peak_measured' in r: print('   fp64 peaks dfma %.2f dmma %.2f' % (r['fp64_dfma_peak_measured'], r['fp64_dmma_peak_measured']))
    except Exception as e:
        print(f, 'FAILED', e)
PY
