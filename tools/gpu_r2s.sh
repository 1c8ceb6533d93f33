#!/bin/bash
mkdir -p gpurun_out
echo "== pytest parity + tcgen05 (-x)"
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tcgen05.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_par.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_par.log | tail -8 | cut -c1-300
echo "== bench benzene 512 walkers"
timeout 600 python bench.py --walkers 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_benzene_512.json 2> gpurun_out/bench_benzene_512.err
echo "rc=$?"; cut -c1-330 gpurun_out/bench_benzene_512.json; tail -2 gpurun_out/bench_benzene_512.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_bz32.csv python bench.py --walkers 32 --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > /dev/null 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_bz32.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=i;break
hd=rows[h]; kn=hd.index('Kernel Name'); mv=hd.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[h+1:]:
    if len(r)>mv:
        try: v=float(r[mv].replace(',',''))
        except: continue
        agg[r[kn].split('(')[0][:50]].append(v)
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:9]: print(f'  {sum(v)/len(v)/1e6:8.3f} ms avg x{len(v):3d}  {k}')
PY
