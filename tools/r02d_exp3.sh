for e in "" "FB200_MEGA_ROWS=0 FB200_TC_MIN=0"; do
  echo "== env: $e"
  env $e python bench.py --no-cpu-baseline --no-extras --sweep 32 --steps 30 2>/dev/null | python tools/show_bench.py /dev/stdin | grep -E "sweep|incumbent B=32"
done
