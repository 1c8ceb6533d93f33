#!/bin/bash
mkdir -p gpurun_out
echo "== ncu full: non-trunk kernels of one benzene E_loc (32 walkers)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"slater_fwd2|attn_fl_f32|slater_kernel|embed_fwd|finalize_kernel" -s 12 -c 8 -o gpurun_out/prof_rest python bench.py --walkers 32 --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_rest.log 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_rest.ncu-rep; tail -3 gpurun_out/ncu_rest.log
