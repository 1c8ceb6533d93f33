#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tcgen05 with CTA pairs"; DQMC_GEMM_2CTA=1 timeout 300 python -m pytest tests/test_gpu_tcgen05.py -x -q -m gpu > gpurun_out/pytest_tc2.log 2>&1; rc=$?; echo "rc=$rc"; tail -15 gpurun_out/pytest_tc2.log | cut -c1-300
if [ $rc -eq 0 ]; then
echo "== bench lih 2cta"; DQMC_GEMM_2CTA=1 timeout 300 python bench.py --workload lih_psiformer --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lih_2cta.json 2> gpurun_out/bench_lih_2cta.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_lih_2cta.json; tail -3 gpurun_out/bench_lih_2cta.err
echo "== bench benzene 1024 2cta"; DQMC_GEMM_2CTA=1 timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --walkers 1024 --equil-sweeps 2 > gpurun_out/bench_benzene_1024_2cta.json 2> gpurun_out/bench_benzene_1024_2cta.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_benzene_1024_2cta.json; tail -3 gpurun_out/bench_benzene_1024_2cta.err
echo "== pytest fp32 parity with CTA pairs"; DQMC_GEMM_2CTA=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fp32 or tensor_core or full_size" > gpurun_out/pytest_par2.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_par2.log | cut -c1-300
fi
echo "== pytest overlap"; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "overlap" > gpurun_out/pytest_ov.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_ov.log | cut -c1-300
