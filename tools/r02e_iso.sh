# r02e: isolates the three tail changes.  base = r02d build; vA = + smalln_bwd warp turns; vB = vA + red.global adds;
# new = vB + finer split of the deferred FFMA weight gradients.
D=multimodal-model-skin-lesion-classifier_b200/build/ab
B="python bench.py --no-cpu-baseline --no-incumbent --no-extras --steps 40"
for i in 1 2; do
  for wl in cfg3a cfg2; do
    for lib in base vA vB new; do
      FB200_LIB=$PWD/$D/libfb200_$lib.so $B --workload $wl --sweep 32,256,1024 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib $wl', round(d['ms_per_step'],4), {k: round(v['ms_per_step'],4) for k,v in (d.get('sweep') or {}).items()})"
    done
  done
done
