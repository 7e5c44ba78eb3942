#!/bin/bash
# usage: tools/sweep.sh "<lanes> <waves per lane> <batch>" ...   (device-resident value only)
for cfg in "$@"; do
  read L W B <<< "$cfg"
  PCOP_LANES=$L PCOP_WAVES_PER_LANE=$W python bench.py --batch $B --cpu-sample 0 --latency-reps 0 --steps 4 --no-kernel-timing > gpurun_out/sweep.json 2> gpurun_out/sweep.err
  echo "lanes $L wpl $W batch $B: $(python tools/bench_summary.py gpurun_out/sweep.json 0 | head -1)"
done
