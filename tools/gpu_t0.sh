OUT=gpurun_out
for wl in t_lin t_bge c3; do
  extra="--no-also"; [ $wl = t_lin ] && extra=""
  DIBS_BENCH_TIMELINE=1 timeout 600 python bench.py --workload $wl --steps $( [ $wl = t_lin ] && echo 200 || echo 100 ) --warmup 5 --no-cpu-baseline $extra 2> $OUT/bench_${wl}_t0.err > $OUT/bench_${wl}_t0.json
  python - <<PY
import json
j=json.loads([l for l in open("$OUT/bench_${wl}_t0.json") if l.startswith("{")][0])
print("$wl steps", j["steps"], "value", round(j["value"],1), "value_from_t0", round(j["value_from_t0"],1), "e2e", round(j["e2e"]["value"],1), "l2_resident", round(j["value_l2_resident"],1))
for k,v in (j.get("also") or {}).items(): print("   also", k, "steps", v["steps"], "value", round(v["value"],1), "value_from_t0", round(v["value_from_t0"],1))
PY
done
