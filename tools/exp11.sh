#!/usr/bin/env bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/exp11_all.log
python bench.py --no-cpu-baseline --steps 50 --warmup 10 > gpurun_out/exp11.json 2> gpurun_out/exp11.err
python bench.py --no-cpu-baseline --workload cfg5 --steps 50 --warmup 10 > gpurun_out/exp11_cfg5.json 2> gpurun_out/exp11_cfg5.err
python tools/tc_trace.py 1 0 4096 512 2048 2>&1 | tail -6 > gpurun_out/exp11_trace.log
python tools/tc_trace.py 2 0 4096 512 2048 2>&1 | tail -6 >> gpurun_out/exp11_trace.log
echo done
