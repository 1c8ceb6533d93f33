#!/bin/bash
mkdir -p gpurun_out
for off in 0 1; do
  if [ $off = 1 ]; then export DQMC_ECP_ENV_TABLE_OFF=1; fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_bz32_off$off.csv python bench.py --walkers 32 --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > /dev/null 2>&1
  python - <<PY
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_bz32_off$off.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=i;break
hd=rows[h]; kn=hd.index('Kernel Name'); mv=hd.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[h+1:]:
    if len(r)>mv:
        try: v=float(r[mv].replace(',',''))
        except: continue
        agg[r[kn].split('(')[0][:50]].append(v)
print('table off =', $off)
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:8]: print(f'  {sum(v)/len(v)/1e6:8.3f} ms avg x{len(v):3d}  {k}')
PY
done
