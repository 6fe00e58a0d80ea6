O=gpurun_out
python -m pytest tests/test_gpu_parity.py -q -k "rgatt or cfg5 or cfg4b or small" 2>&1 | tail -3
for wl in cfg4b cfg5; do
  python bench.py --workload $wl --no-cpu-baseline --no-incumbent --no-extras --sweep "" --steps 40 > $O/r02c_wide1_$wl.json 2>/dev/null
  FB200_GRB_WIDE=0 python bench.py --workload $wl --no-cpu-baseline --no-incumbent --no-extras --sweep "" --steps 40 > $O/r02c_wide0_$wl.json 2>/dev/null
done
python - <<'PY'
import json
for wl in ("cfg4b","cfg5"):
    for w in (1,0):
        d=json.load(open(f"gpurun_out/r02c_wide{w}_{wl}.json")); print(wl, "wide",w, d["ms_per_step"], d["dtype"])
PY
