#!/bin/bash
# elect_one() issue paths: tensor-core tests, trunk ablation + trace, bench lines
mkdir -p gpurun_out
echo "== pytest tcgen05 + full-size"
timeout 900 python -m pytest tests/test_gpu_tcgen05.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_tc.log | tail -8 | cut -c1-300
echo "== trunk ablation"
timeout 300 python tools/trunk_ablate.py 2>&1 | tail -12
echo "== trace"; DQMC_TRUNK_TRACE=1 timeout 300 python tools/trunk_trace.py 2>&1 | tail -2
echo "== trace (no MMA / TMA)"; DQMC_TRUNK_ABLATE=15 DQMC_TRUNK_TRACE=1 timeout 300 python tools/trunk_trace.py 2>&1 | tail -2
echo "== plain forward timing"
timeout 300 python tools/prof_fwd.py 3 2>&1 | tail -2
echo "== bench benzene 512 walkers"
timeout 600 python bench.py --walkers 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_benzene_512.json 2> gpurun_out/bench_benzene_512.err
echo "rc=$?"; cut -c1-400 gpurun_out/bench_benzene_512.json; tail -2 gpurun_out/bench_benzene_512.err
echo "== bench lih"; timeout 300 python bench.py --workload lih_psiformer --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lih.json 2> gpurun_out/bench_lih.err
echo "rc=$?"; cut -c1-300 gpurun_out/bench_lih.json; tail -2 gpurun_out/bench_lih.err
echo "== bench lih eval step"; timeout 300 python bench.py --workload lih_eval_step --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lih_eval.json 2> gpurun_out/bench_lih_eval.err
echo "rc=$?"; cut -c1-300 gpurun_out/bench_lih_eval.json; tail -2 gpurun_out/bench_lih_eval.err
echo "== ncu launch list (benzene, 32 walkers)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_benzene32.csv python bench.py --walkers 32 --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_bz32.log 2>&1
echo "rc=$?"
