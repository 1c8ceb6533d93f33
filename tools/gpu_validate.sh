#!/bin/bash
# Hardware validation of the head of the tree (run under gpurun, one B200; `--gpus 2` adds the two-rank leg):
#   tools/gpurun_retry.sh /tmp/validate.out --timeout 2400 -- 'bash tools/gpu_validate.sh'
# 1. the whole `-m gpu` suite WITHOUT -x, 2. smoke(), 3. the default bench line (benzene, 4096 walkers) and the reference arm,
# 4. the other workloads (LiH, LiH evaluation step, N2 FermiNet, cyclobutadiene 2 states), 5. ncu launch list of one
# benzene local-energy step.  Outputs land in gpurun_out/ (copy what is quoted into profiles/).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
echo "== pytest -m gpu (no -x)"
timeout 1800 python -m pytest tests -q -m gpu --durations=8 -p no:cacheprovider > gpurun_out/pytest_gpu_head.log 2>&1
echo "rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu_head.log | tail -30 | cut -c1-300
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -4
echo "== default bench"
( time timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
echo "rc=$?"; cut -c1-600 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
echo "== reference arm"
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real
cut -c1-400 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
for wl in lih_psiformer lih_eval_step n2_ferminet cyclobutadiene_transpsiformer; do
  echo "== bench $wl"
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  echo "rc=$?"; cut -c1-330 gpurun_out/bench_$wl.json; tail -2 gpurun_out/bench_$wl.err
done
echo "== ncu launch list (benzene, 32 walkers, one step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file gpurun_out/launches_benzene32.csv python bench.py --walkers 32 --steps 1 --warmup 3 --no-cpu-baseline --equil-sweeps 0 > gpurun_out/ncu_bz32.log 2>&1
echo "rc=$?"
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  echo "== torchrun N=2 default bench"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
  echo "rc=$?"; tail -1 gpurun_out/bench_n2.json | cut -c1-600
fi
