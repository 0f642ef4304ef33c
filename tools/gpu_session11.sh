#!/bin/bash
TAG=r02k
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | grep real; tail -2 gpurun_out/bench_$TAG.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err ) 2>&1 | grep real
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_driver_args_$TAG.json 2>> gpurun_out/bench_$TAG.err
timeout 600 python tools/parity_report.py 50 2>&1 | grep -E "==|step 1 |step 2 |step 50|mismatch" | cut -c1-700 > gpurun_out/parity_$TAG.log; cp gpurun_out/parity_report.json gpurun_out/parity_$TAG.json
grep -E "==|element-wise" gpurun_out/parity_$TAG.log
timeout 600 python examples/train_fpv_ppo.py --task pos --num-envs 4096 --epochs 40 > gpurun_out/train_pos_4096_native_$TAG.jsonl 2>/dev/null; tail -1 gpurun_out/train_pos_4096_native_$TAG.jsonl | cut -c1-700
PARTS=ppo bash tools/gpu_sanitize.sh $TAG
bash tools/gpu_traffic.sh $TAG
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
