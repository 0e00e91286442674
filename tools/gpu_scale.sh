#!/bin/bash
# N-GPU bench lines only.  usage: gpurun --gpus N -- 'bash tools/gpu_scale.sh tag N [workloads...]'
TAG=${1:-s}; N=${2:-8}; shift; shift
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for wl in "${@:-c2}"; do
  timeout 600 $TR bench.py --gpus $N --workload $wl --steps ${STEPS:-200} --warmup 10 --no-cpu-baseline 2> $OUT/bench_${wl}_n${N}_$TAG.err | tee $OUT/bench_${wl}_n${N}_$TAG.json | cut -c1-200
  grep -v "^\*\*\*\|OMP_NUM\|^$" $OUT/bench_${wl}_n${N}_$TAG.err | tail -3
done
