#!/bin/bash
# Development aid: same-box A/B of build/variants/lib_<name>.so: correctness vs the FFMA path on 256 samples, then cfg2 E_loc timing.
#   scripts/ab_variants.sh base kpack ...
cd "$(dirname "$0")/.."
for v in "$@"; do
  RNNWF_LIB=$PWD/build/variants/lib_$v.so python scripts/ab_time.py 512 256 2>&1 | tail -1
done
for rep in 1 2; do
  for v in "$@"; do
    RNNWF_LIB=$PWD/build/variants/lib_$v.so python scripts/ab_time.py 10000 2>&1 | tail -1
  done
done
