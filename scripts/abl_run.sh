#!/bin/bash
# Development aid: ablations of the chain kernel (tensor pipe alone / gate math alone) with SM clock and power sampled under load.
cd "$(dirname "$0")/.."
for v in "$@"; do
  echo "== $v"
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits -lms 50 > /tmp/clk_$v.log 2>&1 &
  SMI=$!
  RNNWF_LIB=$PWD/build/variants/lib_$v.so python scripts/ab_time.py 10000 2>&1 | tail -1
  kill $SMI
  python - "$v" <<'PY'
import sys, statistics as st
rows = [tuple(float(x) for x in l.split(",")) for l in open(f"/tmp/clk_{sys.argv[1]}.log") if l.count(",") == 1]
busy = [r for r in rows if r[1] > 350]
if busy:
    print(f"   under load ({len(busy)} samples of {len(rows)}): SM clock median {st.median(r[0] for r in busy):.0f} MHz, power median {st.median(r[1] for r in busy):.0f} W, max {max(r[1] for r in busy):.0f} W")
PY
done
