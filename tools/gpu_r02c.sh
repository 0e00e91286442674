#!/bin/bash
# Round 2: fused step (assemble / exp->K / optimizer inside the producing kernels): parity, bench at N=1 and N ranks.
# usage: gpurun --gpus N --timeout 1800 -- 'bash tools/gpu_r02c.sh tag N'
TAG=${1:-r02c}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 | tee $OUT/pytest_gpu_$TAG.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.txt
for wl in t_lin t_bge; do
  echo "== bench $wl N=1"; timeout 600 python bench.py --workload $wl --steps 200 --warmup 10 --no-cpu-baseline --no-also 2> $OUT/bench_${wl}_n1_$TAG.err > $OUT/bench_${wl}_n1_$TAG.json
  tail -2 $OUT/bench_${wl}_n1_$TAG.err
  if [ "$N" -gt 1 ]; then
    echo "== bench $wl N=$N"; timeout 600 $TR bench.py --gpus $N --workload $wl --steps 200 --warmup 10 --no-cpu-baseline --no-also 2> $OUT/bench_${wl}_n${N}_$TAG.err > $OUT/bench_${wl}_n${N}_$TAG.json
    grep -v "^\*\*\*\|OMP_NUM\|^$" $OUT/bench_${wl}_n${N}_$TAG.err | tail -3
  fi
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_*_r02c*.json')):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', round(j['value'], 1), 'us/step', round(j['ms_per_step'] * 1000, 1), 'e2e', round(j['e2e']['value'], 1),
              {k: round(v['us'], 1) for k, v in j['kernels'].items()})
    except Exception as e:
        print(f, 'unreadable', e)
PY
