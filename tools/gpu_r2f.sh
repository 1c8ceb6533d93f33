#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tcgen05 (no -x)"
timeout 900 python -m pytest tests/test_gpu_tcgen05.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error|assert" gpurun_out/pytest_tc.log | tail -8 | cut -c1-300
echo "== trace"; DQMC_TRUNK_TRACE=1 timeout 300 python tools/trunk_trace.py 2>&1 | tail -3
echo "== plain forward timing"
timeout 300 python tools/prof_fwd.py 3 2>&1 | tail -2
echo "== full-size oracle parity test"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "full_size_one_walker or full_psiformer" > gpurun_out/pytest_full.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error|assert|^E " gpurun_out/pytest_full.log | tail -12 | cut -c1-300
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_fwd.csv python tools/prof_fwd.py 2 > gpurun_out/ncu_list.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_fwd.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
last=rows[-6:]
tot=sum(float(r[vi].replace(',','')) for r in last)
for r in last: print(f"{float(r[vi].replace(',',''))/1e6:9.3f} ms {100*float(r[vi].replace(',',''))/tot:5.1f}%  {r[ki][:90]}")
print('total', tot/1e6, 'ms')
PY
