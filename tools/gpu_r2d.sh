#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tcgen05 (no -x)"
timeout 900 python -m pytest tests/test_gpu_tcgen05.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error|assert" gpurun_out/pytest_tc.log | tail -15 | cut -c1-300
echo "== accuracy benzene (trunk on)"; timeout 900 python tools/acc_study.py benzene 256 2>&1 | tail -8
echo "== accuracy benzene (trunk off)"; DQMC_TC_TRUNK=0 timeout 900 python tools/acc_study.py benzene 256 2>&1 | tail -1
