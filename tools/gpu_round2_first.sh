#!/bin/bash
# First GPU call of the next round: re-validate, then the experiments that were prepared without a GPU.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh tag'
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu_$TAG.txt
# (1) 4x4-tile acyclicity kernel at n_vars <= 32 (one warp per sample, ~60 registers) vs the row-per-lane kernel;
#     parity of the variant was confirmed at the end of round 1 (profiles/r01/pytest_gpu_final.txt), only timing is open
for v in 0 1; do
  DIBS_B200_ACYC_TILE=$v timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>/dev/null > $OUT/bench_c2_${TAG}_tile$v.json
  DIBS_B200_ACYC_TILE=$v timeout 300 python -m pytest tests -m gpu -x -q -k "oracle_n_vars_20 or full_steps" 2>&1 | tail -2
done
# (1b) 64x64-tile phi kernel (8 x 4 per thread) vs the 32x64 one: parity (kernel_and_phi, full steps) and timing at C2 / t_lin
DIBS_B200_PHI_TILE=1 timeout 300 python -m pytest tests -m gpu -x -q -k "kernel_and_phi or full_steps or oracle_n_vars_20" 2>&1 | tail -2
for wl in c2 t_lin; do for v in 0 1; do
  DIBS_B200_PHI_TILE=$v timeout 300 python bench.py --workload $wl --steps 200 --warmup 10 --no-cpu-baseline 2>/dev/null > $OUT/bench_${wl}_${TAG}_phitile$v.json
done; done
# (1c) dense LinearGaussian draw phase V2 (per-CTA precomputed probabilities, no integer divisions, paired threefry lanes)
DIBS_B200_DENSE_V2=1 timeout 300 python -m pytest tests -m gpu -x -q -k "oracle_n_vars_20 and lingauss" 2>&1 | tail -2
for v in 0 1; do
  DIBS_B200_DENSE_V2=$v timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null > $OUT/bench_c5_${TAG}_densetile$v.json
done
# (2) where the time goes in the BGe and DenseNN passes now
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mc_bge|k_mc_nn' -s 4 -c 3 -f -o $OUT/prof_bge_$TAG \
    python bench.py --workload t_bge --steps 6 --warmup 3 --no-cpu-baseline > $OUT/ncu_bge_$TAG.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_*_*tile*.json')):
    j = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(j['ms_per_step'] * 1000, 1), 'us', {k: round(v['us'], 1) for k, v in j['kernels'].items()})
PY
