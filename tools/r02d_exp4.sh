timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_primitives.py -q 2>&1 | tail -1
B="python bench.py --no-cpu-baseline --no-incumbent --no-extras --steps 30 --batch 256"
echo "== cfg2 auto"; $B --sweep 32,64,128,384,512,768,1024,2048 2>/dev/null | python tools/show_bench.py /dev/stdin | grep -E "sweep|B=256"
echo "== cfg3a auto"; $B --workload cfg3a --sweep 32,128,512,1024,2048 2>/dev/null | python tools/show_bench.py /dev/stdin | grep -E "sweep|B=256"
echo "== cfg1 auto / half"; $B --workload cfg1 --sweep 32,512,1024 2>/dev/null | python tools/show_bench.py /dev/stdin | grep -E "sweep|B=256"; FB200_TC_CSPLIT_FILL=2 $B --workload cfg1 --sweep 32,512,1024 2>/dev/null | python tools/show_bench.py /dev/stdin | grep -E "sweep|B=256"
