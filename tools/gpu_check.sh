#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench lines, ncu launch list + full capture of the top kernel.
# usage (from the build container): gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_$TAG.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.txt
echo "== bench c2"; timeout 600 python bench.py 2> $OUT/bench_c2_$TAG.err | tee $OUT/bench_c2_$TAG.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 2 2> /dev/null | tee $OUT/bench_ref_c2_$TAG.json
for wl in t_lin t_bge c4 c3; do
  echo "== bench $wl"; timeout 600 python bench.py --workload $wl --steps 100 --warmup 5 --no-cpu-baseline 2> $OUT/bench_${wl}_$TAG.err | tee $OUT/bench_${wl}_$TAG.json
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_c2_$TAG.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
echo "== ncu full (top kernels)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_mc_lingauss|k_acyclic_grad|k_phi_update|k_pair_dist' -s 8 -c 8 \
    -f -o $OUT/prof_c2_$TAG python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT
