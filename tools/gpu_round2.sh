#!/bin/bash
mkdir -p gpurun_out
echo "== tcgen05 tests"; timeout 300 python -m pytest tests/test_gpu_tcgen05.py -x -q > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_tc.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench tc"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench simt"; timeout 600 python bench.py --steps 10 --warmup 3 --gemm-backend simt --no-cpu-baseline > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err; echo "rc=$?"; cat gpurun_out/bench_simt.json
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"; wc -l gpurun_out/launches.csv
