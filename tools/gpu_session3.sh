#!/bin/bash
TAG=r02c
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "compact or checkpoint or graphed or host_buffer" > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
for i in 1 2; do
  timeout 200 python tools/step_windows.py flip 2097152 12 25 > gpurun_out/windows_flip_${TAG}_diffdev_$i.json 2>/dev/null
  TACO_B200_LIB=$PWD/taco_b200/lib/libtaco_b200_mb6.so timeout 200 python tools/step_windows.py flip 2097152 12 25 > gpurun_out/windows_flip_${TAG}_nodiff_$i.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/windows_flip_r02c_*.json')):
    d=json.load(open(f)); print(f, [w['ms_per_step'] for w in d['windows']])
PY
( time timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err ) 2>&1 | grep real
cut -c1-400 gpurun_out/bench_ref_$TAG.json
bash tools/gpu_traffic.sh $TAG
