#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu (no -x)"
timeout 1500 python -m pytest tests -q -m gpu --durations=3 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu.log | tail -20 | cut -c1-300
for am in 1 0; do
  echo "== bench benzene 512 walkers ATTN_MMA=$am"
  DQMC_ATTN_MMA=$am timeout 600 python bench.py --walkers 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_benzene_512_am$am.json 2> gpurun_out/bench_benzene_512_am$am.err
  echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_benzene_512_am$am.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'roof', d['roofline']['achieved'], 'share', d['roofline']['gemm_share_of_step'], 'E', d['energy_mean'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_benzene_512_am$am.err').read()[-1500:])
PY
done
echo "== plain forward timing"; timeout 300 python tools/prof_fwd.py 3 2>&1 | tail -3
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/launches_fwd.csv python tools/prof_fwd.py 2 > gpurun_out/ncu_list.log 2>&1
echo "rc=$?"; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_fwd.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg={}
for r in rows[1:]:
    try: agg[r[ki][:60]]=agg.get(r[ki][:60],0)+float(r[vi].replace(',',''))
    except Exception: pass
tot=sum(agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1]): print(f'{v/1e6:9.3f} ms {100*v/tot:5.1f}%  {k}')
PY
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm3xtf32|mlp_block|attn_fwd_mma|slater_fwd2" -s 11 -c 5 -o gpurun_out/prof_fwd python tools/prof_fwd.py 2 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
