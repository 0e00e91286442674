#!/bin/bash
# parity tests + bench + one ncu --set full capture of selected kernels.
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_prof.sh tag "regex" [workload]'
TAG=${1:-p}; RX=${2:-k_mc}; WL=${3:-c2}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu_$TAG.txt
timeout 600 python bench.py --workload $WL --steps 200 --warmup 10 --no-cpu-baseline 2> $OUT/bench_${WL}_$TAG.err | tee $OUT/bench_${WL}_$TAG.json
tail -5 $OUT/bench_${WL}_$TAG.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s ${NCU_SKIP:-16} -c ${NCU_COUNT:-8} \
    -f -o $OUT/prof_${WL}_$TAG python bench.py --workload $WL --steps 6 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
tail -3 $OUT/ncu_full_$TAG.log
