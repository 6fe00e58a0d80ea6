#!/usr/bin/env bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/exp9_all.log
python bench.py --no-cpu-baseline --steps 50 --warmup 10 > gpurun_out/exp9.json 2> gpurun_out/exp9.err
python tools/tc_trace.py 1 0 4096 512 2048 2>&1 | tail -14 > gpurun_out/exp9_trace.log
python tools/tc_trace.py 1 2 512 2048 4096 2>&1 | tail -3 >> gpurun_out/exp9_trace.log
echo done
