#!/bin/bash
# N independent single-GPU processes (no torch.distributed, no NCCL) running tools/wave_ab.py side by side: separates what
# the box shares (host cores, PCIe, memory) from what the multi-GPU bench adds.  usage: tools/n8_probe.sh N [reps]
N=${1:-8}; REPS=${2:-20}
NCPU=$(nproc); PER=$((NCPU / N))
run() {  # $1 = label, $2 = pin (0/1), rest = env
  local label=$1 pin=$2; shift 2
  for i in $(seq 0 $((N - 1))); do
    if [ "$pin" = 1 ]; then PRE="taskset -c $((i * PER))-$((i * PER + PER - 1))"; else PRE=""; fi
    env CUDA_VISIBLE_DEVICES=$i "$@" $PRE python tools/wave_ab.py 1024 $REPS > gpurun_out/n8p_${label}_$i.log 2>&1 &
  done
  wait
  echo "== $label"; for i in $(seq 0 $((N - 1))); do tail -1 gpurun_out/n8p_${label}_$i.log; done
}
run host_nopin 0
run host_pin 1
run dev_pin 1 WAVE_AB_DEVICE=1
run dev_nopin 0 WAVE_AB_DEVICE=1
run host_pin_trace 1 PCOP_TRACE=1
nvidia-smi --query-gpu=index,clocks.sm,power.draw,pcie.link.gen.current,pcie.link.width.current --format=csv
lscpu | grep -i "model name\|socket\|numa\|^cpu(s)"
nvidia-smi topo -m | head -14
