#!/bin/bash
# N-GPU pass.  usage: gpurun --gpus N -- 'bash tools/gpu_scale.sh tag N [workloads...]'   (TESTS=1: multi-GPU parity tests first)
TAG=${1:-s}; N=${2:-8}; shift; shift
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "${TESTS:-0}" = "1" ]; then
  timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q ${TEST_K:+-k "$TEST_K"} 2>&1 | tail -4 | tee $OUT/pytest_mgpu_n${N}_$TAG.txt
fi
for wl in "${@:-t_lin}"; do
  DIBS_BENCH_TIMELINE=1 timeout 600 $TR bench.py --gpus $N --workload $wl --steps ${STEPS:-200} --warmup 10 --no-cpu-baseline --no-also \
      2> $OUT/bench_${wl}_n${N}_$TAG.err > $OUT/bench_${wl}_n${N}_$TAG.json
  python - <<PY
import json
try:
    j = json.loads([l for l in open("$OUT/bench_${wl}_n${N}_$TAG.json") if l.startswith("{")][0])
    print("$wl N=$N", round(j["value"], 1), "steps/s", round(j["ms_per_step"] * 1e3, 1), "us  e2e", round(j["e2e"]["value"], 1), j["config"].get("exchange"))
    print("   kernels", {k: v["us"] for k, v in j["kernels"].items()})
    print("   timeline", j.get("timeline_end_us"))
except Exception as e:
    print("no bench line:", e)
PY
  grep -v "^\*\*\*\|OMP_NUM\|^$" $OUT/bench_${wl}_n${N}_$TAG.err | tail -3
done
