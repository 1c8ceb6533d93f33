#!/bin/bash
# trunk v2 (tcgen05 attention) on hardware
mkdir -p gpurun_out
echo "== pytest tcgen05 (no -x)"
timeout 900 python -m pytest tests/test_gpu_tcgen05.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error|assert" gpurun_out/pytest_tc.log | tail -15 | cut -c1-300
echo "== plain forward timing (trunk on / off)"
timeout 300 python tools/prof_fwd.py 3 2>&1 | tail -2
DQMC_TC_TRUNK=0 timeout 300 python tools/prof_fwd.py 3 2>&1 | tail -1
echo "== bench benzene 512 walkers"
timeout 600 python bench.py --walkers 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_benzene_512.json 2> gpurun_out/bench_benzene_512.err
echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_benzene_512.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'roof', d['roofline']['achieved'], 'share', d['roofline']['gemm_share_of_step'], 'launches', d['roofline']['gemm_launches_per_step'], 'E', d['energy_mean'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_benzene_512.err').read()[-1500:])
PY
echo "== ncu launch list (plain forward, benzene, 86400 walkers)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_fwd.csv python tools/prof_fwd.py 2 > gpurun_out/ncu_list.log 2>&1
echo "rc=$?"; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_fwd.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
last=rows[-6:]
tot=sum(float(r[vi].replace(',','')) for r in last)
for r in last: print(f"{float(r[vi].replace(',',''))/1e6:9.3f} ms {100*float(r[vi].replace(',',''))/tot:5.1f}%  {r[ki][:90]}")
print('total', tot/1e6, 'ms')
PY
echo "== ncu full (trunk kernel, 17760 walkers)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trunk_f16" -s 1 -c 1 -o gpurun_out/prof_trunk2 python tools/prof_fwd.py 2 17760 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full.log
