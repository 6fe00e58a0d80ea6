"""Print the fields of a bench.py JSON line that matter when iterating (python tools/show_bench.py file.json ...)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as exc:
        print(path, "unreadable:", exc)
        continue
    print(f"== {path}")
    cfg = d.get("config", {})
    print(f"  {cfg.get('mechanism')} B={cfg.get('per_gpu_batch')} x{d.get('n_gpus')} {d.get('dtype')}: {d.get('ms_per_step'):.4f} ms/step, "
          f"value {d.get('value'):.4g} {d.get('unit')}, e2e {d.get('e2e', {}).get('value'):.4g}, launches {d.get('gpu_launches')}")
    r = d.get("roofline") or {}
    if r:
        print(f"  roofline: {r.get('achieved'):.1f}/{r.get('peak'):.1f} {r.get('unit')} frac {r.get('frac'):.3f}; step frac {r.get('step', {}).get('frac'):.3f}; gemm ms {r.get('gemm_ms_per_step')}")
    for k, v in (d.get("sweep") or {}).items():
        print(f"  sweep B={k}: {v['ms_per_step'] * 1e3:.1f} us/step, roofline frac {v.get('roofline_frac', float('nan')):.3f}")
    for k, v in (d.get("incumbent") or {}).items():
        print(f"  incumbent B={k}: " + ", ".join(f"{kk} {vv:.3f}" for kk, vv in v.items() if isinstance(vv, float)))
    for key in ("dp_check", "cpu_baseline", "clocks"):
        if d.get(key):
            print(f"  {key}: {d[key]}")
