#!/usr/bin/env bash
cd "$(dirname "$0")/.."
for i in 1 2 3; do python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/exp5_pytest$i.log; done
FB200_PDL=1 python bench.py --no-cpu-baseline --steps 50 --warmup 10 > gpurun_out/exp5_pdl1.json 2> gpurun_out/exp5_pdl1.err
echo done
