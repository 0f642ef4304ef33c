#!/bin/bash
# One GPU session: parity tests, bench, ncu launch list + full captures of the step kernel and the actor kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [skip-tests]
TAG=${1:-r01}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  timeout 600 python tools/parity_report.py 50 2>&1 | grep -E "==|step 1 |step 2 |step 50|mismatch" | cut -c1-700 > gpurun_out/parity_$TAG.log
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
  tail -5 gpurun_out/pytest_gpu_$TAG.log
fi
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --fast-fp --no-cpu-baseline --no-actor > gpurun_out/bench_fast_$TAG.json 2>> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_fast_$TAG.json
BENCH_SMALL="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-small"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $BENCH_SMALL > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fpv_step_kernel -s 3 -c 2 -f -o gpurun_out/prof_step_$TAG $BENCH_SMALL --no-actor > gpurun_out/ncu_full_step_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:actor_tc_kernel -s 2 -c 2 -f -o gpurun_out/prof_actor_$TAG $BENCH_SMALL > gpurun_out/ncu_full_actor_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:critic_tc_kernel -s 2 -c 2 -f -o gpurun_out/prof_critic_$TAG $BENCH_SMALL > gpurun_out/ncu_full_critic_$TAG.log 2>&1
ls -la gpurun_out | tail -14
# BASELINE configs 4 / 5 per-GPU shards (steady-state windows) and the actor kernel alone at two sizes
timeout 300 python tools/step_windows.py rotate 524288 8 25 > gpurun_out/windows_rotate_$TAG.json 2>/dev/null
timeout 300 python tools/step_windows.py mix 2097152 8 25 --dr > gpurun_out/windows_mixdr_$TAG.json 2>/dev/null
timeout 300 python tools/step_windows.py flip 2097152 12 25 > gpurun_out/windows_flip_$TAG.json 2>/dev/null
(timeout 100 python tools/actor_bench.py 262144; timeout 100 python tools/actor_bench.py 2097152) > gpurun_out/actor_bench_$TAG.json 2>/dev/null
TACO_ACTOR_TIMELINE=gpurun_out/actor_tl_$TAG.bin timeout 100 python tools/actor_bench.py 4096 256,256,256 2 > /dev/null 2>&1
(timeout 100 python tools/critic_bench.py 262144; timeout 100 python tools/critic_bench.py 2097152 64 256,256,256 10) > gpurun_out/critic_bench_$TAG.json 2>/dev/null
TACO_CRITIC_TIMELINE=gpurun_out/critic_tl_$TAG.bin timeout 100 python tools/critic_bench.py 4096 64 256,256,256 2 > /dev/null 2>&1
timeout 200 python tools/host_step_probe.py flip 2097152 30 > gpurun_out/host_step_probe_$TAG.json 2>/dev/null
