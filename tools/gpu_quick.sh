#!/bin/bash
# quick GPU pass: parity tests + default bench line.  usage: gpurun --timeout 900 -- 'bash tools/gpu_quick.sh tag [workloads...]'
TAG=${1:-q}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu_$TAG.txt
for wl in "${@:-c2}"; do
  timeout 600 python bench.py --workload $wl --steps 200 --warmup 10 --no-cpu-baseline 2> $OUT/bench_${wl}_$TAG.err | tee $OUT/bench_${wl}_$TAG.json
  tail -5 $OUT/bench_${wl}_$TAG.err
done
