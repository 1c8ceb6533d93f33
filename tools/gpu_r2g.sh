#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tcgen05 (no -x)"
timeout 900 python -m pytest tests/test_gpu_tcgen05.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error|assert" gpurun_out/pytest_tc.log | tail -8 | cut -c1-300
echo "== trace"; DQMC_TRUNK_TRACE=1 timeout 300 python tools/trunk_trace.py 2>&1 | tail -2
echo "== plain forward timing"
timeout 300 python tools/prof_fwd.py 3 2>&1 | tail -2
echo "== overlap test (cyclobutadiene)"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "excited_state_overlap" > gpurun_out/pytest_ov.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error|assert|^E " gpurun_out/pytest_ov.log | tail -8 | cut -c1-300
echo "== bench benzene 512 walkers"
timeout 600 python bench.py --walkers 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_benzene_512.json 2> gpurun_out/bench_benzene_512.err
echo "rc=$?"; cut -c1-400 gpurun_out/bench_benzene_512.json; tail -2 gpurun_out/bench_benzene_512.err
echo "== bench lih"; timeout 300 python bench.py --workload lih_psiformer --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lih.json 2> gpurun_out/bench_lih.err
echo "rc=$?"; cut -c1-300 gpurun_out/bench_lih.json; tail -2 gpurun_out/bench_lih.err
echo "== bench cyclobutadiene 2 states"; timeout 600 python bench.py --workload cyclobutadiene_transpsiformer --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cbd.json 2> gpurun_out/bench_cbd.err
echo "rc=$?"; cut -c1-700 gpurun_out/bench_cbd.json; tail -3 gpurun_out/bench_cbd.err
