#!/bin/bash
# two-GPU check exactly as the driver launches it
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
echo "== torchrun N=2 default bench"
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err ) 2>&1 | grep real
echo "rc=$?"; tail -1 gpurun_out/bench_n2.json | cut -c1-1500; grep -E "NCCL INFO (comm|Connected|NVLS)|nranks|Error|error" gpurun_out/bench_n2.err | head -8
echo "== torchrun N=2 reference arm"
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err ) 2>&1 | grep real
tail -1 gpurun_out/bench_ref_n2.json | cut -c1-400
echo "== torchrun N=2 lih"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload lih_psiformer --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lih_n2.json 2> gpurun_out/bench_lih_n2.err
tail -1 gpurun_out/bench_lih_n2.json | cut -c1-400
