#!/bin/bash
# multi-GPU pass: consistency vs single GPU, then bench lines at N ranks.  usage: gpurun --gpus N -- 'bash tools/gpu_mgpu.sh tag N [workloads...]'
TAG=${1:-m}; N=${2:-2}; shift; shift
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python tools/mgpu_check.py 2>&1 | tail -5
timeout 300 $TR tools/mgpu_check.py 2>&1 | tail -8
python tools/mgpu_check.py --compare $N 2>&1 | tee $OUT/mgpu_compare_${TAG}.txt
for wl in "${@:-c2}"; do
  timeout 600 python bench.py --workload $wl --steps 200 --warmup 10 --no-cpu-baseline 2> $OUT/bench_${wl}_n1_$TAG.err | tee $OUT/bench_${wl}_n1_$TAG.json | cut -c1-200
  timeout 600 $TR bench.py --gpus $N --workload $wl --steps 200 --warmup 10 --no-cpu-baseline 2> $OUT/bench_${wl}_n${N}_$TAG.err | tee $OUT/bench_${wl}_n${N}_$TAG.json | cut -c1-200
  tail -3 $OUT/bench_${wl}_n${N}_$TAG.err
done
