export DIBS_BENCH_TIMELINE=1
run() { tag=$1; wl=$2; shift; shift; env "$@" python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu-baseline --no-also 2>/dev/null > gpurun_out/x_$tag.json; }
run base_c2 c2 A=1
run prio_c2 c2 DIBS_X_PRIO=1
run base_tl t_lin A=1
run prio_tl t_lin DIBS_X_PRIO=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/x_*_c2.json")+glob.glob("gpurun_out/x_*_tl.json")):
    try: j=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,e); continue
    print(f, round(j["value"]), round(j["ms_per_step"]*1000,1), "hot", round(1e6/j["value_l2_resident"],1), j.get("timeline_end_us"))
PY
