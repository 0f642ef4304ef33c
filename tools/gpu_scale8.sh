#!/bin/bash
# 8-GPU session (gpurun --gpus 8): default bench + BASELINE configs 4 (rotate, 524288 envs/GPU) and 5 (mix + per-env DR, 2 Mi envs/GPU)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --no-small --no-actor > gpurun_out/bench_${N}gpu_flip.json 2> gpurun_out/bench_${N}gpu.err; echo "flip rc=$?"
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 200 --warmup 100 --task rotate --envs-per-gpu 524288 --no-cpu-baseline --no-small --no-actor --no-e2e > gpurun_out/bench_${N}gpu_config4_rotate.json 2>> gpurun_out/bench_${N}gpu.err; echo "config4 rc=$?"
timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 200 --warmup 100 --task mix --dr --no-cpu-baseline --no-small --no-actor --no-e2e > gpurun_out/bench_${N}gpu_config5_mixdr.json 2>> gpurun_out/bench_${N}gpu.err; echo "config5 rc=$?"
cat gpurun_out/bench_${N}gpu_*.json | cut -c1-600; tail -5 gpurun_out/bench_${N}gpu.err
