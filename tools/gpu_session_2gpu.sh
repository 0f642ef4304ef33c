#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/ppo_native_ddp_check.py > gpurun_out/ppo_native_ddp_2gpu_r02j.json 2> gpurun_out/ppo_native_ddp_2gpu_r02j.err; echo "ddp rc=$?"; cat gpurun_out/ppo_native_ddp_2gpu_r02j.json; tail -3 gpurun_out/ppo_native_ddp_2gpu_r02j.err
timeout 600 $TR bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/bench_2gpu_r02j.json 2> gpurun_out/bench_2gpu_r02j.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_2gpu_r02j.json; tail -3 gpurun_out/bench_2gpu_r02j.err
timeout 600 $TR examples/train_fpv_ppo.py --task mix --num-envs 65536 --horizon 32 --epochs 12 > gpurun_out/train_mix_2gpu_native_r02j.jsonl 2> gpurun_out/train_mix_2gpu_native_r02j.err; echo "train rc=$?"; tail -1 gpurun_out/train_mix_2gpu_native_r02j.jsonl | cut -c1-800; tail -3 gpurun_out/train_mix_2gpu_native_r02j.err
