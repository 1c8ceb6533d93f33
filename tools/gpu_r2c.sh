#!/bin/bash
# whole-trunk kernel on hardware: tests, A/B bench lines, plain-forward launch list, ncu full capture of the trunk kernel
mkdir -p gpurun_out
echo "== pytest tcgen05"
timeout 900 python -m pytest tests/test_gpu_tcgen05.py -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error|error|assert" gpurun_out/pytest_tc.log | tail -15 | cut -c1-300
echo "== plain forward timing (trunk on / off)"
timeout 300 python tools/prof_fwd.py 3 2>&1 | tail -3
DQMC_TC_TRUNK=0 timeout 300 python tools/prof_fwd.py 3 2>&1 | tail -1
for tr in 1 0; do
  echo "== bench benzene 512 walkers TRUNK=$tr"
  DQMC_TC_TRUNK=$tr timeout 600 python bench.py --walkers 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_benzene_512_t$tr.json 2> gpurun_out/bench_benzene_512_t$tr.err
  echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_benzene_512_t$tr.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'roof', d['roofline']['achieved'], 'share', d['roofline']['gemm_share_of_step'], 'launches', d['roofline']['gemm_launches_per_step'], 'E', d['energy_mean'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_benzene_512_t$tr.err').read()[-1500:])
PY
done
echo "== bench lih"; timeout 300 python bench.py --workload lih_psiformer --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lih.json 2> gpurun_out/bench_lih.err
echo "rc=$?"; cut -c1-250 gpurun_out/bench_lih.json
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trunk_f16" -s 1 -c 1 -o gpurun_out/prof_trunk python tools/prof_fwd.py 2 17760 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
