#!/bin/bash
# Development aid: per-kernel times (ncu launch list) of gradient calls at 10^4 samples for build/variants/lib_<name>.so.
cd "$(dirname "$0")/.."
for v in "$@"; do
  echo "== $v (ncu launch list, gradient calls at 10^4 samples)"
  RNNWF_LIB=$PWD/build/variants/lib_$v.so ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/grad_launches_1e4_$v.csv python scripts/grad_once.py 10000 > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/grad_launches_1e4_$v.csv 6
done
