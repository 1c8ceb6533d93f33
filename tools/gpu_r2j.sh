#!/bin/bash
# TS-form trunk (A operand in tensor memory): tests, ablation / trace A-B against the SS form, bench lines
mkdir -p gpurun_out
echo "== pytest tcgen05"
timeout 900 python -m pytest tests/test_gpu_tcgen05.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_tc.log | tail -8 | cut -c1-300
echo "== trunk ablation TS"
timeout 300 python tools/trunk_ablate.py 2>&1 | tail -12
echo "== trunk ablation SS"
DQMC_TC_TRUNK_TS=0 timeout 300 python tools/trunk_ablate.py 2>&1 | tail -12 | head -3
echo "== trace TS"; DQMC_TRUNK_TRACE=1 timeout 300 python tools/trunk_trace.py 2>&1 | tail -2
echo "== plain forward timing"
timeout 300 python tools/prof_fwd.py 3 2>&1 | tail -2
echo "== full-size parity tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "benzene or full_size" > gpurun_out/pytest_full.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_full.log | tail -8 | cut -c1-300
echo "== bench benzene 512 walkers"
timeout 600 python bench.py --walkers 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_benzene_512.json 2> gpurun_out/bench_benzene_512.err
echo "rc=$?"; cut -c1-400 gpurun_out/bench_benzene_512.json; tail -2 gpurun_out/bench_benzene_512.err
