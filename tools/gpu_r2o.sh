#!/bin/bash
mkdir -p gpurun_out
echo "== pytest parity (slater paths) "
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tcgen05.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_par.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_par.log | tail -8 | cut -c1-300
echo "== plain forward timing"
timeout 300 python tools/prof_fwd.py 3 2>&1 | tail -2
echo "== bench benzene 512 walkers"
timeout 600 python bench.py --walkers 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_benzene_512.json 2> gpurun_out/bench_benzene_512.err
echo "rc=$?"; cut -c1-330 gpurun_out/bench_benzene_512.json; tail -2 gpurun_out/bench_benzene_512.err
echo "== ncu launch list (benzene, 512 walkers, 1 step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 3000 --csv --log-file gpurun_out/launches_benzene512.csv python bench.py --walkers 512 --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_bz512.log 2>&1
echo "rc=$?"
