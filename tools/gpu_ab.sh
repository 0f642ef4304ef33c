#!/bin/bash
# usage (under gpurun): bash tools/gpu_ab.sh <tag> <rounds> <lib> ...
TAG=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/ab_$TAG.txt
timeout 1500 python tools/step_ab.py "$@" >> gpurun_out/ab_$TAG.txt 2>&1
cat gpurun_out/ab_$TAG.txt
