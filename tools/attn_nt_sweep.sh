#!/bin/bash
mkdir -p gpurun_out
echo "== tcgen05 tests"
timeout 1500 python -m pytest tests/test_gpu_tcgen05.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_new.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|^E  " gpurun_out/pytest_new.log | tail -12 | cut -c1-300
echo "== trunk ablation"
timeout 300 python tools/trunk_ablate.py 2>&1 | tail -12 | head -2
echo "benzene: $(timeout 300 python bench.py --walkers 512 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['energy_mean'])")"
