#!/bin/bash
# Final single-GPU pass of a round: parity tests, smoke, default bench (north-star shape + also-block + CPU baseline),
# reference arm, ncu launch list and one full capture of the step kernels (traffic table + summaries for profiles/).
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_final.sh tag'
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $OUT/pytest_gpu_$TAG.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.txt
echo "== bench default"; DIBS_BENCH_TIMELINE=1 timeout 900 python bench.py --steps 200 --warmup 10 2> $OUT/bench_t_lin_$TAG.err > $OUT/bench_t_lin_$TAG.json; tail -2 $OUT/bench_t_lin_$TAG.err; cut -c1-300 $OUT/bench_t_lin_$TAG.json
echo "== bench driver flags"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2> /dev/null > $OUT/bench_t_lin_${TAG}_driverflags.json; cut -c1-200 $OUT/bench_t_lin_${TAG}_driverflags.json
echo "== bench reference"; timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2> /dev/null > $OUT/bench_ref_t_lin_$TAG.json; cut -c1-300 $OUT/bench_ref_t_lin_$TAG.json
for wl in c2 t_bge c3 c4; do
  echo "== bench $wl"; DIBS_BENCH_TIMELINE=1 timeout 600 python bench.py --workload $wl --steps 100 --warmup 5 --no-cpu-baseline --no-also 2> $OUT/bench_${wl}_$TAG.err > $OUT/bench_${wl}_$TAG.json; cut -c1-160 $OUT/bench_${wl}_$TAG.json
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_t_lin_$TAG.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-also > $OUT/ncu_bench_$TAG.log 2>&1
echo "== ncu full (step kernels)"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_mc_lin_qr|k_acyclic_rows|k_phi|k_pair_dist|k_pair_finish|k_prologue|k_edge_probs' -s 14 -c 9 \
    -f -o $OUT/prof_t_lin_$TAG python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-also > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log
if [ -f _exp/libdibs_b200_trace.so ]; then
  echo "== phi pipeline trace (debug build of the library, tools/phi_trace.py)"
  timeout 200 python tools/phi_trace.py _exp/libdibs_b200_trace.so t_lin > $OUT/phi_trace_t_lin_$TAG.txt 2>&1; tail -14 $OUT/phi_trace_t_lin_$TAG.txt
fi
ls -la $OUT | tail -12
