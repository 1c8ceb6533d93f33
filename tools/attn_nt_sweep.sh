#!/bin/bash
# A/B of the forward-Laplacian attention variants (development aid)
mkdir -p gpurun_out
echo "== pytest (transpsiformer, overlap, fp32 tolerance, tcgen05 parity)"
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -k "transpsiformer or overlap or fp32 or full" > gpurun_out/pytest_par.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_par.log | tail -8 | cut -c1-300
for m in 1 0; do
  echo "cyclobutadiene MMA=$m: $(DQMC_ATTN_FL_MMA=$m timeout 600 python bench.py --workload cyclobutadiene_transpsiformer --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['energy_mean'])")"
done
echo "benzene: $(timeout 300 python bench.py --walkers 512 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['energy_mean'])")"
