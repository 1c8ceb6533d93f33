#!/bin/bash
# First GPU call of the next round: everything written after round 1's GPU budget ran out was verified on the CPU
# emulator only.  Run the hardware-verified files first, then the new file WITHOUT -x so that every
# failure is listed, then smoke() and the short bench lines.  Usage: gpurun --timeout 1800 -- 'bash tools/gpu_round2_first.sh'
mkdir -p gpurun_out
echo "== pytest hardware-verified files"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tcgen05.py -q -m gpu --durations=3 > gpurun_out/pytest_verified.log 2>&1
echo "rc=$?"; tail -6 gpurun_out/pytest_verified.log | cut -c1-300
echo "== pytest next rows / reference fixtures (no -x)"
timeout 1200 python -m pytest tests/test_gpu_z_next_rows.py -q -m gpu --durations=5 > gpurun_out/pytest_next_rows.log 2>&1
echo "rc=$?"; grep -E "FAILED|ERROR|passed|failed" gpurun_out/pytest_next_rows.log | tail -30 | cut -c1-300
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/smoke.log | cut -c1-300
echo "== bench lih"; timeout 300 python bench.py --workload lih_psiformer --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lih.json 2> gpurun_out/bench_lih.err
echo "rc=$?"; cut -c1-600 gpurun_out/bench_lih.json
echo "== bench benzene 1024 walkers"; timeout 600 python bench.py --walkers 1024 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_benzene_1024.json 2> gpurun_out/bench_benzene_1024.err
echo "rc=$?"; cut -c1-600 gpurun_out/bench_benzene_1024.json
