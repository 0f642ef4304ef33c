T=r01i
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$T.log
timeout 600 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"
(timeout 100 python tools/actor_bench.py 262144; timeout 100 python tools/actor_bench.py 2097152) > gpurun_out/actor_bench_$T.json 2>/dev/null
(timeout 100 python tools/critic_bench.py 262144; timeout 100 python tools/critic_bench.py 2097152 64 256,256,256 10) > gpurun_out/critic_bench_$T.json 2>/dev/null
BENCH_SMALL="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-small"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_$T.csv $BENCH_SMALL > gpurun_out/ncu_launch_$T.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:actor_tc_kernel -s 2 -c 2 -f -o gpurun_out/prof_actor_$T $BENCH_SMALL > gpurun_out/ncu_full_actor_$T.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:critic_tc_kernel -s 2 -c 2 -f -o gpurun_out/prof_critic_$T $BENCH_SMALL > gpurun_out/ncu_full_critic_$T.log 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$T.json"))
print(d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["config3_mix_actor"]["value"], d["config3_mix_actor"]["actor_tc_ms"], d["config3_mix_actor"]["actor_roofline"]["frac"], d["config3_mix_actor"]["with_critic"]["value"], d["config3_mix_actor"]["with_critic"]["critic_roofline"]["frac"])
PY
cat gpurun_out/actor_bench_$T.json
