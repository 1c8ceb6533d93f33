#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tcgen05 first"; timeout 300 python -m pytest tests/test_gpu_tcgen05.py -x -q -m gpu > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_tc.log
echo "== pytest all"; timeout 700 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -9 gpurun_out/pytest_gpu.log
echo "== bench benzene 1024"; timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --walkers 1024 --equil-sweeps 2 > gpurun_out/bench_benzene_1024.json 2> gpurun_out/bench_benzene_1024.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_benzene_1024.json; tail -3 gpurun_out/bench_benzene_1024.err
echo "== bench lih"; timeout 300 python bench.py --workload lih_psiformer --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lih.json 2> gpurun_out/bench_lih.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_lih.json
echo "== ncu launches benzene"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_benzene.csv python bench.py --walkers 32 --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_benzene.log 2>&1; echo "ncu rc=$?"
echo "== ncu full"; timeout 400 ncu --set full --clock-control none -k regex:"gemm3xtf32|slater_fwd2|attn_fwd_f32" -s 41 -c 6 -o /tmp/prof_fwd python bench.py --walkers 8 --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_fwd.ncu-rep --page raw --csv > gpurun_out/prof_fwd_raw.csv 2>/dev/null; ls -la gpurun_out/ | head -20
