#!/bin/bash
# compute-sanitizer over every kernel family (small sizes); summaries -> gpurun_out/sanitizer/
cd "$(dirname "$0")/.."
out=gpurun_out/sanitizer; mkdir -p $out
tools=${TOOLS:-"memcheck racecheck synccheck"}
cases=${CASES:-"tc16p tc16p1 cplx grad ffma mdrnn misc"}
for tool in $tools; do
  for c in $cases; do
    log=$out/${tool}_$c.log
    timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_cases.py $c > $log 2>&1
    echo "$tool $c rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)" | tee -a $out/summary.txt
  done
done
