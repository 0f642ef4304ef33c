#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r02f.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_r02f.log
timeout 300 python tools/ppo_native_time.py 5 > gpurun_out/ppo_time_r02f.json 2>gpurun_out/ppo_time_r02f.err; cat gpurun_out/ppo_time_r02f.json
timeout 600 python examples/train_fpv_ppo.py --task pos --num-envs 4096 --epochs 40 > gpurun_out/train_pos_4096_native_r02f.jsonl 2> gpurun_out/train_pos_4096_native_r02f.err; echo "train rc=$?"
tail -2 gpurun_out/train_pos_4096_native_r02f.jsonl | cut -c1-900; tail -3 gpurun_out/train_pos_4096_native_r02f.err
timeout 600 python examples/train_fpv_ppo.py --task pos --num-envs 4096 --epochs 12 --update torch > gpurun_out/train_pos_4096_torch_r02f.jsonl 2>/dev/null
tail -1 gpurun_out/train_pos_4096_torch_r02f.jsonl | cut -c1-700
timeout 600 python tools/ppo_native_check.py --no-time > gpurun_out/ppo_native_check_r02f.log 2>&1; tail -3 gpurun_out/ppo_native_check_r02f.log | cut -c1-1200
