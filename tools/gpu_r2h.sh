#!/bin/bash
# HEAD validation: whole -m gpu suite (no -x), smoke(), default bench line (4096 benzene walkers), reference arm.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
echo "== pytest -m gpu (no -x)"
timeout 1500 python -m pytest tests -q -m gpu --durations=8 -p no:cacheprovider > gpurun_out/pytest_gpu_head.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu_head.log | tail -30 | cut -c1-300
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -4
echo "== default bench"
( time timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
echo "rc=$?"; cut -c1-3000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
echo "== reference arm"
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real
cut -c1-1200 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
echo "== trunk ablation"
timeout 300 python tools/trunk_ablate.py 2>&1 | tail -12
