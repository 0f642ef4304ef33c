#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ppo_native_check.py > gpurun_out/ppo_native_check.log 2>&1; echo "ppo check rc=$?"
tail -12 gpurun_out/ppo_native_check.log | cut -c1-1500
for i in 1 2; do
  timeout 200 python tools/step_windows.py flip 2097152 10 25 > gpurun_out/windows_flip_r02d_cur_$i.json 2>/dev/null
  ( cd build_mbr01 && timeout 200 python tools/step_windows.py flip 2097152 10 25 > ../gpurun_out/windows_flip_r02d_r01src_$i.json 2>/dev/null )
  TACO_B200_LIB=$PWD/taco_b200/lib/libtaco_b200_mb6.so timeout 200 python tools/step_windows.py flip 2097152 10 25 > gpurun_out/windows_flip_r02d_nodiff_$i.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/windows_flip_r02d_*.json')):
    try:
        d=json.load(open(f)); print(f, [w['ms_per_step'] for w in d['windows']])
    except Exception as e: print(f, 'ERR', e)
PY
