#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"slater_fwd2" -s 4 -c 1 -o gpurun_out/prof_slater python bench.py --walkers 32 --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_slater.log 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_slater.ncu-rep
