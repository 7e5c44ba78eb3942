#!/bin/bash
# usage: tools/sweep2.sh "<env assignments> | <batch>" ...
for cfg in "$@"; do
  envs="${cfg%%|*}"; B="${cfg##*|}"
  env $envs python bench.py --batch $B --cpu-sample 0 --latency-reps 0 --steps 4 --no-kernel-timing > gpurun_out/sweep.json 2> gpurun_out/sweep.err
  echo "$envs batch $B: $(python tools/bench_summary.py gpurun_out/sweep.json 0 | head -1)"
done
