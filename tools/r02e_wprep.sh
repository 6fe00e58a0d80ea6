# r02e: bf16 weight copies made on the side lane (FB200_WPREP_SIDE=1; measured mixed, off by default) against the caller's stream (FB200_WPREP_SIDE=0), same box; bf16 tests
B="python bench.py --no-cpu-baseline --no-incumbent --no-extras --steps 40"
for i in 1 2; do for s in 0 1; do for wl in cfg4b cfg5; do
  FB200_WPREP_SIDE=$s $B --workload $wl --sweep 256 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('side=$s $wl', round(d['ms_per_step'],4), {k: round(v['ms_per_step'],4) for k,v in (d.get('sweep') or {}).items()})"
done; done; done
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "bf16" 2>&1 | tail -2
