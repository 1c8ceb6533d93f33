#!/bin/bash
# A/B of the forward-Laplacian attention variants (development aid)
mkdir -p gpurun_out
echo "== pytest parity (-x)"
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tcgen05.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_par.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_par.log | tail -8 | cut -c1-300
for cfg in "1 4" "1 2" "1 3" "0 4"; do
  set -- $cfg
  echo "MMA=$1 TB=$2: $(DQMC_ATTN_FL_MMA=$1 DQMC_ATTN_TB=$2 timeout 300 python bench.py --walkers 512 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['energy_mean'])")"
done
for m in 1 0; do
  echo "LiH MMA=$m: $(DQMC_ATTN_FL_MMA=$m timeout 300 python bench.py --workload lih_psiformer --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['energy_mean'])")"
done
