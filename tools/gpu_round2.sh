#!/bin/bash
# One GPU session producing the round-2 evidence under gpurun_out/ (copy what is to be kept into profiles/).
# usage (under gpurun): bash tools/gpu_round2.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_driver_args_$TAG.json 2>> gpurun_out/bench_$TAG.err
timeout 600 python tools/parity_report.py 50 2>&1 | grep -E "==|step 1 |step 2 |step 50|mismatch" | cut -c1-700 > gpurun_out/parity_$TAG.log; cp gpurun_out/parity_report.json gpurun_out/parity_$TAG.json
timeout 600 python examples/train_fpv_ppo.py --task pos --num-envs 4096 --epochs 40 > gpurun_out/train_pos_4096_native_$TAG.jsonl 2>/dev/null
timeout 600 python tools/run_reference_ppo.py --task flip --epochs 20 --horizon 64 --out gpurun_out/reference_ppo_flip_$TAG.jsonl > /dev/null 2>&1
timeout 600 python tools/ppo_native_check.py > gpurun_out/ppo_native_check_$TAG.log 2>&1; cp gpurun_out/ppo_native_check.json gpurun_out/ppo_native_check_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_ppo_$TAG.csv python tools/ppo_native_time.py 1 > /dev/null 2>&1
bash tools/gpu_sanitize.sh $TAG
bash tools/gpu_traffic.sh $TAG
# multi-GPU (gpurun --gpus N):
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N --steps 200 --warmup 20
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ppo_native_ddp_check.py
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 examples/train_fpv_ppo.py --task mix --num-envs 65536 --horizon 32 --epochs 12
