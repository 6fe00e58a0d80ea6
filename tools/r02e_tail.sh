# r02e: same-box A/B of the r02d build (base) against the build with the smalln_bwd warp-turn reduction, red.global adds and the
# finer split of the deferred FFMA weight gradient (new); then the whole GPU suite on the new build.
D=multimodal-model-skin-lesion-classifier_b200/build/ab
B="python bench.py --no-cpu-baseline --no-incumbent --no-extras --steps 40"
for i in 1 2; do
  for lib in base new; do
    for wl in cfg2 cfg3a; do
      FB200_LIB=$PWD/$D/libfb200_$lib.so $B --workload $wl --sweep 32,64,128,256,512,1024 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib $wl', round(d['ms_per_step'],4), {k: round(v['ms_per_step'],4) for k,v in (d.get('sweep') or {}).items()})"
    done
  done
done
echo "== eager routes at B=32 (incumbent block)"
for lib in base new; do FB200_LIB=$PWD/$D/libfb200_$lib.so python bench.py --no-cpu-baseline --no-extras --steps 20 --sweep "" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', {k: {a: round(b,4) for a,b in v.items() if a != 'note'} for k,v in d['incumbent'].items()})"; done
echo "== GPU suite on the new build"
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -2
