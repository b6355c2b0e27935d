#!/bin/bash
# usage: scripts/clock_watch.sh <command...>  — samples SM clock / power / throttle reasons every 100 ms while the command runs
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active,clocks_throttle_reasons.sw_power_cap,clocks_throttle_reasons.hw_slowdown,temperature.gpu --format=csv,noheader -lms 100 > /tmp/clk.log 2>&1 &
SMI=$!
"$@"
kill $SMI
sort /tmp/clk.log | uniq -c | sort -rn | head -12
