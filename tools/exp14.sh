#!/usr/bin/env bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/exp14_all.log
python bench.py --no-cpu-baseline --steps 50 --warmup 10 --sweep 32,128,256,1024 > gpurun_out/exp14.json 2> gpurun_out/exp14.err
FB200_LANES=0 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --batch 32 > gpurun_out/exp14_b32_l0.json 2> /dev/null
python bench.py --no-cpu-baseline --steps 20 --warmup 5 --batch 32 --workload cfg1 > gpurun_out/exp14_cfg1_b32.json 2> /dev/null
echo done
