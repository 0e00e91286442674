#!/bin/bash
# Round 2, first GPU call: new parity cases, the one-GPU two-rank exchange test, default bench (north-star shape),
# timing of the three prepared experiment kernels, ncu launch list + full capture at the target shape.
TAG=${1:-r02a}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee $OUT/pytest_gpu_$TAG.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.txt
echo "== bench default"; timeout 900 python bench.py --steps 200 --warmup 10 2> $OUT/bench_t_lin_$TAG.err | tee $OUT/bench_t_lin_$TAG.json | cut -c1-400
tail -3 $OUT/bench_t_lin_$TAG.err
for v in 1; do
  echo "== ACYC_TILE=$v"; DIBS_B200_ACYC_TILE=$v timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-also 2>/dev/null > $OUT/bench_t_lin_${TAG}_acyctile$v.json
  echo "== PHI_TILE=$v"; DIBS_B200_PHI_TILE=$v timeout 300 python -m pytest tests -m gpu -x -q -k "kernel_and_phi or full_steps" 2>&1 | tail -2
  DIBS_B200_PHI_TILE=$v timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-also 2>/dev/null > $OUT/bench_t_lin_${TAG}_phitile$v.json
done
echo "== DENSE_V2"; DIBS_B200_DENSE_V2=1 timeout 300 python -m pytest tests -m gpu -x -q -k "oracle and lingauss" 2>&1 | tail -2
for v in 0 1; do
  DIBS_B200_DENSE_V2=$v timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null > $OUT/bench_c5_${TAG}_densev2$v.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_*r02a*.json')):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j['ms_per_step'] * 1000, 1), 'us', {k: round(v['us'], 1) for k, v in j['kernels'].items()})
    except Exception as e:
        print(f, 'unreadable', e)
PY
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_t_lin_$TAG.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-also > $OUT/ncu_bench_$TAG.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_mc_lin_qr|k_acyclic_rows|k_phi_partial|k_opt_update|k_assemble_grad|k_pair_dist|k_pair_finish' -s 16 -c 8 \
    -f -o $OUT/prof_t_lin_$TAG python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-also > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log
ls -la $OUT | tail -20
