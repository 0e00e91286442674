#!/bin/bash
# quick GPU pass: parity tests + bench lines.  usage: gpurun --timeout 900 -- 'bash tools/gpu_quick.sh tag [workloads...]'
TAG=${1:-q}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu_$TAG.txt
for wl in "${@:-c2}"; do
  DIBS_BENCH_TIMELINE=1 timeout 600 python bench.py --workload $wl --steps ${STEPS:-200} --warmup 10 --no-cpu-baseline --no-also 2> $OUT/bench_${wl}_$TAG.err > $OUT/bench_${wl}_$TAG.json
  python - <<PY
import json
j=json.loads([l for l in open("$OUT/bench_${wl}_$TAG.json") if l.startswith("{")][0])
print("$wl", round(j["value"],1), "steps/s", round(j["ms_per_step"]*1e3,1), "us  e2e", round(j["e2e"]["value"],1))
print("   kernels", {k:v["us"] for k,v in j["kernels"].items()})
print("   timeline", j.get("timeline_end_us"))
PY
  tail -3 $OUT/bench_${wl}_$TAG.err
done
