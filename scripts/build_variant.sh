#!/bin/bash
# Development aid: build build/variants/lib_<name>.so with extra nvcc flags for gru.cu (A/B timing through RNNWF_LIB=...).
#   scripts/build_variant.sh <name> [-DRNNWF_GATES=1 ...]
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/variants build/var_obj
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-extended-lambda --expt-relaxed-constexpr \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -c rnnwavefunctions_b200/csrc/gru.cu -o build/var_obj/gru_$name.o 2> build/var_obj/gru_$name.log
nvcc -shared -o build/variants/lib_$name.so build/obj/capi.o build/var_obj/gru_$name.o build/obj/mdrnn.o build/obj/misc.o build/obj/umma_selftest.o \
  -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
echo build/variants/lib_$name.so
