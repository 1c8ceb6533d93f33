#!/bin/bash
mkdir -p gpurun_out
echo "== pytest all"; timeout 600 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -14 gpurun_out/pytest_gpu.log
echo "== bench lih"; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench.json
echo "== bench benzene 512"; timeout 400 python bench.py --workload benzene_psiformer --steps 2 --warmup 3 --no-cpu-baseline --walkers 512 --equil-sweeps 2 > gpurun_out/bench_benzene_512.json 2> gpurun_out/bench_benzene_512.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_benzene_512.json; tail -3 gpurun_out/bench_benzene_512.err
echo "== bench n2 ferminet"; timeout 300 python bench.py --workload n2_ferminet --steps 5 --warmup 3 --no-cpu-baseline --equil-sweeps 2 > gpurun_out/bench_n2f.json 2> gpurun_out/bench_n2f.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_n2f.json; tail -3 gpurun_out/bench_n2f.err
echo "== ncu launches benzene"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_benzene.csv python bench.py --workload benzene_psiformer --walkers 32 --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_benzene.log 2>&1; echo "ncu rc=$?"
