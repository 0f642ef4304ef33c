#!/bin/bash
# 8-GPU session, default bench only (e2e with NUMA-bound pinned buffers)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --no-small --no-actor 2> gpurun_out/bench_${N}gpu_b.err | grep '^{' > gpurun_out/bench_${N}gpu_flip_b.json; echo "flip rc=$?"
TACO_HOST_MODE_NOTE=copy timeout 300 $TR --master-port 29522 bench.py --gpus $N --steps 50 --warmup 10 --no-cpu-baseline --no-small --no-actor --no-numa 2>> gpurun_out/bench_${N}gpu_b.err | grep '^{' > gpurun_out/bench_${N}gpu_flip_nonuma.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*gpu_flip_*.json")):
    try:
        d = json.load(open(f)); print(f, d["value"], d["e2e"], d["config"].get("rank0_numa_bound_cpus"))
    except Exception as e:
        print(f, "ERR", e)
PY
nvidia-smi topo -m 2>/dev/null | head -14; tail -3 gpurun_out/bench_${N}gpu_b.err
