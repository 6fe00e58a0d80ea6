#!/usr/bin/env bash
cd "$(dirname "$0")/.."
for b in 32 64 128; do for e in simt tc; do
  python bench.py --no-cpu-baseline --steps 20 --warmup 5 --batch $b --engine $e > gpurun_out/exp15_${b}_${e}.json 2> /dev/null
done; done
echo done
