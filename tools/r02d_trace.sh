timeout 300 python -m pytest tests/test_gpu_primitives.py -q -k tcgen05 2>&1 | tail -1
timeout 200 python tools/r02d_gemm_sweep.py 2>&1 | grep "bad cases"
for c in 8 0; do echo "== FB200_TC_CSPLIT=$c  fp32 M=256 N=512 K=512 layout 0"; FB200_TC_CSPLIT=$c python tools/tc_trace.py 1 0 256 512 512 2>/dev/null | grep -E "entry->setup|cluster epi"; done
echo "== fp32 M=64 N=512 K=512 (BN=64? S=4)"; python tools/tc_trace.py 1 0 64 512 512 2>/dev/null | grep -E "entry->setup|cluster epi"
python bench.py --no-cpu-baseline --no-incumbent --no-extras --sweep "64,128,256,512" --steps 30 2>/dev/null | python tools/show_bench.py /dev/stdin | grep -v incumbent
