#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tcgen05 (1-CTA)"; timeout 300 python -m pytest tests/test_gpu_tcgen05.py -x -q -m gpu > gpurun_out/pytest_tc1.log 2>&1; rc=$?; echo "rc=$rc"; tail -3 gpurun_out/pytest_tc1.log | cut -c1-300
echo "== pytest tcgen05 (CTA pairs)"; DQMC_GEMM_2CTA=1 timeout 300 python -m pytest tests/test_gpu_tcgen05.py -x -q -m gpu > gpurun_out/pytest_tc2.log 2>&1; rc2=$?; echo "rc=$rc2"; tail -3 gpurun_out/pytest_tc2.log | cut -c1-300
echo "== bench lih"; timeout 300 python bench.py --workload lih_psiformer --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lih.json 2> gpurun_out/bench_lih.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_lih.json
echo "== bench lih 2cta"; DQMC_GEMM_2CTA=1 timeout 300 python bench.py --workload lih_psiformer --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lih_2cta.json 2> gpurun_out/bench_lih_2cta.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_lih_2cta.json
echo "== bench benzene 1024"; timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --walkers 1024 --equil-sweeps 2 > gpurun_out/bench_benzene_1024.json 2> gpurun_out/bench_benzene_1024.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_benzene_1024.json
echo "== bench benzene 1024 2cta"; DQMC_GEMM_2CTA=1 timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --walkers 1024 --equil-sweeps 2 > gpurun_out/bench_benzene_1024_2cta.json 2> gpurun_out/bench_benzene_1024_2cta.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_benzene_1024_2cta.json
echo "== pytest transpsiformer"; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "transpsiformer or overlap" > gpurun_out/pytest_tp.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_tp.log | cut -c1-300
