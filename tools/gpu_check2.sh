#!/bin/bash
# 2-GPU check: torchrun bench lines (strong scaling: the global batch is split over the ranks), NCCL set-up logged.
mkdir -p gpurun_out
for wl in lih_psiformer; do
  echo "== bench $wl N=2"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${wl}_n2.json 2> gpurun_out/bench_${wl}_n2.err
  echo "rc=$?"; tail -1 gpurun_out/bench_${wl}_n2.json | cut -c1-700; grep -i "nranks\|NVLS\|error" gpurun_out/bench_${wl}_n2.err | head -5
done
echo "== bench benzene 512 walkers N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --walkers 512 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_benzene_512_n2.json 2> gpurun_out/bench_benzene_512_n2.err
echo "rc=$?"; tail -1 gpurun_out/bench_benzene_512_n2.json | cut -c1-700
echo "== bench lih N=1 (same build)"
timeout 300 python bench.py --workload lih_psiformer --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lih_n1.json 2> gpurun_out/bench_lih_n1.err
echo "rc=$?"; cut -c1-300 gpurun_out/bench_lih_n1.json
