#!/bin/bash
# Quick GPU session: the gpu-marked tests, then a short bench.  usage (under gpurun): bash tools/gpu_check.sh <tag> [pytest -k expr]
TAG=${1:-chk}
KEXPR=${2:-}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/pytest_gpu_$TAG.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
fi
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -40 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
