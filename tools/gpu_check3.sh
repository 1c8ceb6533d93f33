#!/bin/bash
# tensor-core kernels (half-operand GEMM, fused MLP block) on hardware + A/B bench lines
mkdir -p gpurun_out
echo "== pytest tcgen05"
timeout 900 python -m pytest tests/test_gpu_tcgen05.py -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error|error" gpurun_out/pytest_tc.log | tail -15 | cut -c1-300
for cfg in "1 1" "1 0" "0 0"; do
  set -- $cfg
  echo "== bench benzene 512 walkers F16=$1 FUSE=$2"
  DQMC_TC_F16=$1 DQMC_TC_FUSE_MLP=$2 timeout 600 python bench.py --walkers 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_benzene_512_f$1_m$2.json 2> gpurun_out/bench_benzene_512_f$1_m$2.err
  echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_benzene_512_f$1_m$2.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'roof', d['roofline']['achieved'], 'share', d['roofline']['gemm_share_of_step'], 'launches', d['roofline']['gemm_launches_per_step'], 'E', d['energy_mean'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_benzene_512_f$1_m$2.err').read()[-1500:])
PY
done
echo "== bench lih"; timeout 300 python bench.py --workload lih_psiformer --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lih.json 2> gpurun_out/bench_lih.err
echo "rc=$?"; cut -c1-250 gpurun_out/bench_lih.json
