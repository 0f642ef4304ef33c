#!/bin/bash
# compute-sanitizer over every hand-written kernel (tools/sanitize_workload.py).  usage (under gpurun): bash tools/gpu_sanitize.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  for part in ${PARTS:-step nets gae ppo}; do
    timeout 900 $CS --tool $tool --print-limit 20 python tools/sanitize_workload.py $part > gpurun_out/sanitizer_${TAG}_${tool}_${part}.txt 2>&1
    echo "$tool $part rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${TAG}_${tool}_${part}.txt | tail -1)"
  done
done
