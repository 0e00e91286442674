#!/bin/bash
# ncu --set full of ONE k_mc_bge launch at t = 0 and ONE at t = 100 on the same fresh particles (tools/t_dependence.py runs:
# 2 warm-up launches, 3 at t = 0, 3 at t = 100 fresh, ...).   usage: gpurun -- 'bash tools/gpu_bge_ncu.sh t_bge c3'
OUT=gpurun_out; mkdir -p $OUT
WLS="$@"
for wl in $WLS; do
  for pt in "t0 2" "t100 5"; do
    set -- $pt; name=$1; skip=$2
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_mc_bge -s $skip -c 1 -f -o $OUT/prof_bge_${wl}_$name \
        python tools/t_dependence.py $wl > $OUT/ncu_bge_${wl}_$name.log 2>&1
    tail -1 $OUT/ncu_bge_${wl}_$name.log | cut -c1-120
  done
done
ls -la $OUT/prof_bge_* 2>/dev/null
