#!/bin/bash
mkdir -p gpurun_out
echo "== accuracy LiH"; timeout 600 python tools/acc_study.py LiH 1024 2>&1 | tail -8
echo "== accuracy benzene"; timeout 900 python tools/acc_study.py benzene 256 2>&1 | tail -8
echo "== ncu launch list (plain forward, benzene, 86400 walkers)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_fwd.csv python tools/prof_fwd.py 2 > gpurun_out/ncu_list.log 2>&1
echo "rc=$?"; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_fwd.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
# last forward only = the last 16 launches
last=rows[-16:]
tot=sum(float(r[vi].replace(',','')) for r in last)
for r in last: print(f"{float(r[vi].replace(',',''))/1e6:9.3f} ms {100*float(r[vi].replace(',',''))/tot:5.1f}%  {r[ki][:90]}")
print('total', tot/1e6, 'ms')
PY
