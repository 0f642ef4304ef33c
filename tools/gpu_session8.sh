#!/bin/bash
mkdir -p gpurun_out
# launch order inside one optimiser step (33 GEMMs): 0-3 actor fwd, 4-8 LSTM fwd, 9-12 critic fwd, 13.. backward
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 67 -c 24 -f -o gpurun_out/prof_gemm_r02g python tools/ppo_native_time.py 1 --no-lip > gpurun_out/ncu_gemm_r02g.log 2>&1
ls -la gpurun_out/prof_gemm_r02g.ncu-rep
ncu -i gpurun_out/prof_gemm_r02g.ncu-rep --page raw --csv > gpurun_out/prof_gemm_r02g_raw.csv 2>/dev/null
