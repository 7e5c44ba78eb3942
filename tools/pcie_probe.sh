#!/bin/bash
N=${1:-8}
T=$(python -c "import time; print(time.time() + 25)")
for i in $(seq 0 $((N - 1))); do CUDA_VISIBLE_DEVICES=$i python tools/pcie_probe.py $T > gpurun_out/pcie_$i.log 2>&1 & done
wait
for i in $(seq 0 $((N - 1))); do cat gpurun_out/pcie_$i.log; done
