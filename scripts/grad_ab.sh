#!/bin/bash
# Development aid: gradient-stage timing of build/variants/lib_<name>.so at cfg2 + per-kernel times from an ncu launch list.
cd "$(dirname "$0")/.."
for v in "$@"; do
  echo "== $v"
  RNNWF_LIB=$PWD/build/variants/lib_$v.so python scripts/grad_time.py 10000 2>&1 | tail -2
done
for v in "$@"; do
  echo "== $v (ncu launch list, one gradient call)"
  RNNWF_LIB=$PWD/build/variants/lib_$v.so ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/grad_launches_$v.csv python scripts/grad_once.py > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/grad_launches_$v.csv 6
done
