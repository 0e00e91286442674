#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench lines, ncu launch list + full capture of the step kernels.
# usage (from the build container): gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_$TAG.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.txt
echo "== bench c2 (default flags)"; timeout 600 python bench.py 2> $OUT/bench_c2_$TAG.err | tee $OUT/bench_c2_$TAG.json | cut -c1-300
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 2 2> /dev/null | tee $OUT/bench_ref_c2_$TAG.json | cut -c1-300
for wl in t_lin t_bge c4 c3; do
  echo "== bench $wl"; timeout 600 python bench.py --workload $wl --steps 200 --warmup 10 --no-cpu-baseline 2> $OUT/bench_${wl}_$TAG.err | tee $OUT/bench_${wl}_$TAG.json | cut -c1-200
done
echo "== bench c5 (all 4096 particles on one GPU)"; timeout 600 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline 2> $OUT/bench_c5_$TAG.err | tee $OUT/bench_c5_$TAG.json | cut -c1-200
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_c2_$TAG.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
echo "== ncu full (step kernels)"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_mc_lin_qr|k_acyclic_rows|k_phi_partial|k_opt_update|k_assemble_grad|k_pair_dist|k_pair_finish' -s 16 -c 8 \
    -f -o $OUT/prof_c2_$TAG python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT | tail -30
