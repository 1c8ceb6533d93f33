#!/bin/bash
# Round-2 first call at HEAD: the whole `-m gpu` suite WITHOUT -x, then A/B bench lines of the new tensor-core paths.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
echo "== pytest -m gpu (no -x)"
timeout 1500 python -m pytest tests -q -m gpu --durations=5 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu.log | tail -30 | cut -c1-300
for cfg in "1 1 1" "1 1 0" "1 0 0" "0 0 0"; do
  set -- $cfg
  echo "== bench benzene 512 walkers F16=$1 FUSE=$2 ATTN_MMA=$3"
  DQMC_TC_F16=$1 DQMC_TC_FUSE_MLP=$2 DQMC_ATTN_MMA=$3 timeout 600 python bench.py --walkers 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_benzene_512_f$1_m$2_a$3.json 2> gpurun_out/bench_benzene_512_f$1_m$2_a$3.err
  echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_benzene_512_f$1_m$2_a$3.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'roof', d['roofline']['achieved'], 'share', d['roofline']['gemm_share_of_step'], 'launches', d['roofline']['gemm_launches_per_step'], 'E', d['energy_mean'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_benzene_512_f$1_m$2_a$3.err').read()[-1500:])
PY
done
echo "== bench lih"; timeout 300 python bench.py --workload lih_psiformer --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lih.json 2> gpurun_out/bench_lih.err
echo "rc=$?"; cut -c1-250 gpurun_out/bench_lih.json
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
