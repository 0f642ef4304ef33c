#!/bin/bash
# ncu --set full of single launches of the update's GEMM kernel (run under gpurun):  bash tools/gpu_ncu_gemm.sh <tag> <skip> [<skip> ...]
# launch order inside one optimiser step: 5 LSTM steps, 4 critic MLP forward, 4 actor forward, then the backward chain
TAG=$1; shift
mkdir -p gpurun_out
for SKIP in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip $SKIP --launch-count 1 \
    -o gpurun_out/ncu_gemm_${TAG}_skip$SKIP -f python tools/ppo_native_time.py 1 > /dev/null 2>&1
  ncu -i gpurun_out/ncu_gemm_${TAG}_skip$SKIP.ncu-rep --page raw --csv > gpurun_out/ncu_gemm_${TAG}_skip$SKIP.csv 2>/dev/null
done
