# r02e: classifier-head backward, column-owner kernel (new default) against the warp-per-row kernel (FB200_SMALLN_ROWS=1), same box
B="python bench.py --no-cpu-baseline --no-incumbent --no-extras --steps 40"
for i in 1 2; do
  for leg in 1 0; do
    for wl in cfg2 cfg3a; do
      FB200_SMALLN_ROWS=$leg $B --workload $wl --sweep 32,256,1024 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('rows=$leg $wl', round(d['ms_per_step'],4), {k: round(v['ms_per_step'],4) for k,v in (d.get('sweep') or {}).items()})"
    done
  done
done
for leg in 1 0; do FB200_SMALLN_ROWS=$leg $B --workload cfg5 --sweep 256 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('rows=$leg cfg5', round(d['ms_per_step'],4), {k: round(v['ms_per_step'],4) for k,v in (d.get('sweep') or {}).items()})"; done
echo "== parity"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_primitives.py tests/test_gpu_optimizer.py -q -m gpu -x 2>&1 | tail -2
