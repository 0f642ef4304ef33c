#!/bin/bash
# ncu --set full capture of the step kernel (flip, 2 Mi envs) -> gpurun_out/traffic_bytes_per_env.json (copy it to profiles/).
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 20 --no-cpu-baseline --no-e2e --no-small --no-actor"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fpv_step_kernel -s 22 -c 2 -f -o gpurun_out/prof_step_$TAG $B > gpurun_out/ncu_full_step_$TAG.log 2>&1
python tools/update_traffic.py gpurun_out/prof_step_$TAG.ncu-rep flip 2097152 ncu_full_fpv_step_$TAG
python tools/ncu_summary.py gpurun_out/prof_step_$TAG.ncu-rep fpv_step_kernel > gpurun_out/ncu_full_fpv_step_$TAG.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fpv_step_kernel -s 22 -c 2 -f -o gpurun_out/prof_step_mixdr_$TAG $B --task mix --dr > gpurun_out/ncu_full_step_mixdr_$TAG.log 2>&1
python tools/update_traffic.py gpurun_out/prof_step_mixdr_$TAG.ncu-rep mix_dr 2097152 ncu_full_fpv_step_$TAG
python tools/ncu_summary.py gpurun_out/prof_step_mixdr_$TAG.ncu-rep fpv_step_kernel > gpurun_out/ncu_full_fpv_step_mixdr_$TAG.txt 2>&1
# gpurun merges at most 64 MiB back: keep the flip report (source page for tools/ncu_source_lines.py), drop the mix + DR one
python tools/ncu_source_lines.py gpurun_out/prof_step_$TAG.ncu-rep 60 > gpurun_out/ncu_source_lines_fpv_step_$TAG.txt
rm -f gpurun_out/prof_step_mixdr_$TAG.ncu-rep
cat gpurun_out/traffic_bytes_per_env.json
