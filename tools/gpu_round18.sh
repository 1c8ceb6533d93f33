#!/bin/bash
mkdir -p gpurun_out
echo "== pytest all"; timeout 900 python -m pytest tests -x -q -m gpu --durations=3 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
