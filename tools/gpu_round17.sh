#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tcgen05 (1-CTA)"; timeout 300 python -m pytest tests/test_gpu_tcgen05.py -x -q -m gpu > gpurun_out/pytest_tc1.log 2>&1; rc=$?; echo "rc=$rc"; tail -3 gpurun_out/pytest_tc1.log | cut -c1-300
echo "== bench lih"; timeout 300 python bench.py --workload lih_psiformer --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lih.json 2> gpurun_out/bench_lih.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_lih.json
echo "== bench lih 2cta"; DQMC_GEMM_2CTA=1 timeout 300 python bench.py --workload lih_psiformer --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lih_2cta.json 2> gpurun_out/bench_lih_2cta.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_lih_2cta.json
echo "== bench benzene 1024"; timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --walkers 1024 --equil-sweeps 2 > gpurun_out/bench_benzene_1024.json 2> gpurun_out/bench_benzene_1024.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_benzene_1024.json
