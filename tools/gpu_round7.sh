#!/bin/bash
mkdir -p gpurun_out
echo "== pytest benzene"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k benzene > gpurun_out/pytest_benzene.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_benzene.log
echo "== bench benzene"; timeout 500 python bench.py --workload benzene_psiformer --steps 3 --warmup 3 --no-cpu-baseline --walkers 1024 > gpurun_out/bench_benzene_1024.json 2> gpurun_out/bench_benzene_1024.err; echo "rc=$?"; cat gpurun_out/bench_benzene_1024.json; tail -3 gpurun_out/bench_benzene_1024.err
echo "== bench n2 psiformer"; timeout 300 python bench.py --workload n2_psiformer --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2mol.json 2> gpurun_out/bench_n2mol.err; echo "rc=$?"; cat gpurun_out/bench_n2mol.json; tail -3 gpurun_out/bench_n2mol.err
echo "== ncu launches benzene"; timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_benzene.csv python bench.py --workload benzene_psiformer --walkers 256 --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_benzene.log 2>&1; echo "ncu rc=$?"
