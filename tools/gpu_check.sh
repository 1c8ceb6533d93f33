#!/bin/bash
# One GPU call that validates HEAD on a B200: the whole `-m gpu` suite WITHOUT -x (every failure listed), smoke(), and short
# bench lines.  Usage: gpurun --timeout 2400 -- 'bash tools/gpu_check.sh [quick]'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
echo "== pytest -m gpu (no -x)"
timeout 1500 python -m pytest tests -q -m gpu --durations=8 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu.log | tail -40 | cut -c1-400
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/smoke.log | cut -c1-300
echo "== bench lih"; timeout 300 python bench.py --workload lih_psiformer --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lih.json 2> gpurun_out/bench_lih.err
echo "rc=$?"; cut -c1-900 gpurun_out/bench_lih.json; tail -3 gpurun_out/bench_lih.err
echo "== bench benzene 512 walkers"; timeout 600 python bench.py --walkers 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_benzene_512.json 2> gpurun_out/bench_benzene_512.err
echo "rc=$?"; cut -c1-900 gpurun_out/bench_benzene_512.json; tail -3 gpurun_out/bench_benzene_512.err
