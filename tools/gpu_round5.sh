#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tc+parity"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== bench tc (new attn)"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench tc (generic attn)"; DQMC_ATTN_GENERIC=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_genattn.json 2> gpurun_out/bench_genattn.err; echo "bench rc=$?"; cat gpurun_out/bench_genattn.json
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"; wc -l gpurun_out/launches.csv
echo "== ncu full gemm"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm3xtf32 -s 17 -c 4 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_gemm.log 2>&1; echo "rc=$?"
echo "== ncu full attn+slater"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_fl|slater|tanh_fl" -s 13 -c 4 -o gpurun_out/prof_misc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_misc.log 2>&1; echo "rc=$?"
