timeout 600 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_parity.py -q 2>&1 | tail -2
timeout 200 python tools/r02d_gemm_sweep.py 2>&1 | grep "bad cases"
bash tools/r02d_ab.sh
python bench.py --no-cpu-baseline --no-incumbent --no-extras --sweep "64,128,256,512" --steps 30 2>/dev/null | python tools/show_bench.py /dev/stdin | grep -v incumbent
python bench.py --workload cfg5 --no-cpu-baseline --no-incumbent --no-extras --sweep "32,128" --steps 30 2>/dev/null | python tools/show_bench.py /dev/stdin | grep -v incumbent
