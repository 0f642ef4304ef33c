#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ppo_native_time.py 3 > gpurun_out/ppo_time_r02e.json 2>gpurun_out/ppo_time_r02e.err; cat gpurun_out/ppo_time_r02e.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_ppo_r02e.csv python tools/ppo_native_time.py 1 > gpurun_out/ncu_ppo_r02e.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_ppo_r02e.csv')))
hdr = None; agg = collections.OrderedDict(); tot=0
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum': continue
    name = d['Kernel Name'].split('(')[0][-60:]
    v = float(d['Metric Value'].replace(',', ''))
    unit = d['Metric Unit']
    us = v/1000 if unit in ('ns','nsecond') else (v if unit in ('us','usecond') else v*1000)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; tot += us
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t:10.1f} us  {c:5d} x  {t/c:8.1f} us  {k}")
print('total', tot)
PY
for i in 1 2; do
  timeout 200 python tools/step_windows.py flip 2097152 10 25 > gpurun_out/windows_flip_r02e_cur_$i.json 2>/dev/null
  ( cd build_mbr01 && timeout 200 python tools/step_windows.py flip 2097152 10 25 > ../gpurun_out/windows_flip_r02e_r01src_$i.json 2>/dev/null )
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/windows_flip_r02e_*.json')):
    try:
        d=json.load(open(f)); print(f, [w['ms_per_step'] for w in d['windows']])
    except Exception as e: print(f, 'ERR', e)
PY
timeout 900 python -m pytest tests -m gpu -q -x -k "rollout or graphed or checkpoint or ppo_native" 2>&1 | tail -15
