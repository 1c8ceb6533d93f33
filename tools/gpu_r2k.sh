#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tcgen05"
timeout 900 python -m pytest tests/test_gpu_tcgen05.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_tc.log | tail -8 | cut -c1-300
echo "== trunk ablation TS"
timeout 300 python tools/trunk_ablate.py 2>&1 | tail -12
echo "== trace TS"; DQMC_TRUNK_TRACE=1 timeout 300 python tools/trunk_trace.py 2>&1 | tail -2
echo "== plain forward timing"
timeout 300 python tools/prof_fwd.py 3 2>&1 | tail -2
echo "== ncu full trunk"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trunk_f16 -s 1 -c 1 -o gpurun_out/prof_trunk_ts python tools/prof_fwd.py 2 17760 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_trunk_ts.ncu-rep
