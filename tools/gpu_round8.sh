#!/bin/bash
mkdir -p gpurun_out
echo "== pytest all"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench lih"; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench.json
echo "== bench benzene"; timeout 500 python bench.py --workload benzene_psiformer --steps 3 --warmup 3 --no-cpu-baseline --walkers 1024 > gpurun_out/bench_benzene_1024.json 2> gpurun_out/bench_benzene_1024.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_benzene_1024.json; tail -3 gpurun_out/bench_benzene_1024.err
echo "== bench n2 psiformer"; timeout 300 python bench.py --workload n2_psiformer --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2mol.json 2> gpurun_out/bench_n2mol.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_n2mol.json
echo "== ncu launches benzene"; timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_benzene.csv python bench.py --workload benzene_psiformer --walkers 256 --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_benzene.log 2>&1; echo "ncu rc=$?"
echo "== ncu launches n2"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_n2.csv python bench.py --workload n2_psiformer --walkers 1024 --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_n2.log 2>&1; echo "ncu rc=$?"
