mkdir -p gpurun_out/r2d
python scripts/cfg_once.py cfg3 1280 && python scripts/cfg_once.py cfg4 1280 || exit 1
for c in cfg3 cfg4; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d/${c}_launches.csv python scripts/cfg_once.py $c 1280 > gpurun_out/r2d/${c}_ncu.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:gru_chain_kernel -c 1 -o gpurun_out/r2d/cfg3_chain python scripts/cfg_once.py cfg3 1280 > gpurun_out/r2d/cfg3_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mdrnn -c 6 -o gpurun_out/r2d/cfg4_mdrnn python scripts/cfg_once.py cfg4 1280 > gpurun_out/r2d/cfg4_full.log 2>&1
TOOLS="memcheck" scripts/sanitize_all.sh
