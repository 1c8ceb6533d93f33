#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for tb in 12 6 4 3; do
  echo "== bench TB=$tb"; DQMC_ATTN_TB=$tb timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tb$tb.json 2> gpurun_out/bench_tb$tb.err; echo "rc=$?"
done
echo "== bench nofuse"; DQMC_NO_FUSE_TANH=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nofuse.json 2> gpurun_out/bench_nofuse.err; echo "rc=$?"
echo "== bench default + cpu"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
