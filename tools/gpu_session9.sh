#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ppo_native_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python tools/ppo_native_time.py 5 2>/dev/null
timeout 300 python tools/ppo_native_time.py 5 --no-lip 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_ppo_r02h.csv python tools/ppo_native_time.py 1 > gpurun_out/ncu_ppo_r02h.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/launches_ppo_r02h.csv')))
hdr=None; seq=[]
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None or len(r)!=len(hdr): continue
    d=dict(zip(hdr,r))
    if d.get('Metric Name')!='gpu__time_duration.sum': continue
    name=d['Kernel Name']; v=float(d['Metric Value'].replace(',',''))
    unit=d['Metric Unit']; us = v/1000 if unit.startswith('n') else v
    short = 'gemm' if 'gemm_tc' in name else name.split('(')[0][-22:]
    seq.append((short, us))
idxs=[i for i,(n,_) in enumerate(seq) if 'gather' in n]
i0=idxs[-1]
print(' '.join(f"{n[:6]}:{us:.0f}" for n,us in seq[i0:i0+50]))
print('step total us', sum(us for _,us in seq[i0:i0+50]))
PY
