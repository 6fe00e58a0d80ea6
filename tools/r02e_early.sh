# r02e: tensor memory released before the epilogue stores (early) against the build before it (new); then the GPU suite.
D=multimodal-model-skin-lesion-classifier_b200/build/ab
B="python bench.py --no-cpu-baseline --no-incumbent --no-extras --steps 40"
python tools/tc_handover.py 1 0 16384 512 512
python tools/tc_handover.py 2 0 16384 512 512
for i in 1 2; do
  for lib in new early; do
    FB200_LIB=$PWD/$D/libfb200_$lib.so $B --sweep 32,256,1024 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib cfg2', round(d['ms_per_step'],4), 'gemm', round(d['roofline']['gemm_ms_per_step'],4), {k: round(v['ms_per_step'],4) for k,v in (d.get('sweep') or {}).items()})"
  done
done
for lib in new early; do
  FB200_LIB=$PWD/$D/libfb200_$lib.so $B --workload cfg5 --sweep 256 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib cfg5', round(d['ms_per_step'],4), 'gemm', round(d['roofline']['gemm_ms_per_step'],4), {k: round(v['ms_per_step'],4) for k,v in (d.get('sweep') or {}).items()})"
done
echo "== GPU suite"
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -2
