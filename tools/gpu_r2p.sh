#!/bin/bash
mkdir -p gpurun_out
echo "== pytest ECP / benzene parity"
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_z_next_rows.py -q -m gpu -p no:cacheprovider -k "ecp or benzene or full_size or potentials" > gpurun_out/pytest_ecp.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_ecp.log | tail -8 | cut -c1-300
echo "== bench benzene 512 walkers"
timeout 600 python bench.py --walkers 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_benzene_512.json 2> gpurun_out/bench_benzene_512.err
echo "rc=$?"; cut -c1-330 gpurun_out/bench_benzene_512.json; tail -2 gpurun_out/bench_benzene_512.err
echo "== same, table off"
DQMC_ECP_ENV_TABLE_OFF=1 timeout 600 python bench.py --walkers 512 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-330
