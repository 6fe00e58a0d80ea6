# r02e: same-box A/B of the bf16 tcgen05 GEMM with one CTA per SM and a 6-stage ring (base) against two CTAs per SM and
# 3-stage rings (occ2); FB200_LIB selects the build.  Then the bf16 parity tests on the new build.
D=multimodal-model-skin-lesion-classifier_b200/build/ab
B="python bench.py --no-cpu-baseline --no-incumbent --no-extras --steps 40"
for i in 1 2; do
  for lib in base occ2; do
    for wl in cfg4b cfg5; do
      FB200_LIB=$PWD/$D/libfb200_$lib.so $B --workload $wl --sweep 256,1024 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib $wl', round(d['ms_per_step'],4), 'gemm', round(d['roofline'].get('gemm_ms_per_step'),4), 'frac', round(d['roofline']['frac'],3), {k: round(v['ms_per_step'],4) for k,v in (d.get('sweep') or {}).items()})"
    done
  done
done
echo "== fp32 headline on the new build (must be unchanged)"
for lib in base occ2; do FB200_LIB=$PWD/$D/libfb200_$lib.so $B --sweep "" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib cfg2', round(d['ms_per_step'],4))"; done
echo "== parity on occ2"
FB200_LIB=$PWD/$D/libfb200_occ2.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_primitives.py -q -m gpu 2>&1 | tail -2
