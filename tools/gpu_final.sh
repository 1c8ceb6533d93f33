#!/bin/bash
# end-of-round check as the driver runs it: GPU tests, smoke(), default bench line (with CPU baseline)
mkdir -p gpurun_out
echo "== pytest all"; timeout 900 python -m pytest tests -x -q -m gpu --durations=3 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log | cut -c1-300
echo "== bench default"; timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; cut -c1-2500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
