# same-box A/B of builds of the library at the headline batch (FB200_LIB selects the .so; "" = the in-tree build)
D=multimodal-model-skin-lesion-classifier_b200/build/ab
for i in 1 2 3; do
  for lib in "" "$D/libfb200_nocl.so" "$D/libfb200_r02c.so"; do
    FB200_LIB=$lib python bench.py --no-cpu-baseline --no-incumbent --no-extras --sweep "" --steps 60 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('${lib:-new}'[-14:], d['ms_per_step'], d['roofline'].get('gemm_ms_per_step'), d['roofline']['achieved'])"
  done
done
