#!/bin/bash
mkdir -p gpurun_out
echo "== ncu launches benzene"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_benzene.csv python bench.py --walkers 32 --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_benzene.log 2>&1; echo "ncu rc=$?"
echo "== ncu full"; timeout 500 ncu --set full --clock-control none --import-source on -k regex:"gemm3xtf32|slater_fwd2|attn_fwd_f32|embed_fwd" -s 40 -c 23 -o /tmp/prof_fwd python bench.py --walkers 8 --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_fwd.ncu-rep --page raw --csv > gpurun_out/prof_fwd_raw.csv 2>/dev/null; ls -la /tmp/prof_fwd.ncu-rep gpurun_out/
