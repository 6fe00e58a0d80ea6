B="python bench.py --no-cpu-baseline --no-incumbent --no-extras --steps 30 --batch 256"
echo "== default"; $B --sweep 8,16,32,33,48 2>/dev/null | python tools/show_bench.py /dev/stdin | grep sweep
echo "== FB200_MEGA_ROWS=0 FB200_TC_MIN=0 (tcgen05 + cluster split-K at every batch)"; FB200_MEGA_ROWS=0 FB200_TC_MIN=0 $B --sweep 8,16,32,33,48 2>/dev/null | python tools/show_bench.py /dev/stdin | grep sweep
echo "== cfg3a default / tc"; $B --workload cfg3a --sweep 32 2>/dev/null | python tools/show_bench.py /dev/stdin | grep sweep;  FB200_MEGA_ROWS=0 FB200_TC_MIN=0 $B --workload cfg3a --sweep 32 2>/dev/null | python tools/show_bench.py /dev/stdin | grep sweep
