#!/usr/bin/env bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/exp7_all.log
FB200_LANES=1 python bench.py --no-cpu-baseline --steps 50 --warmup 10 > gpurun_out/exp7_l1.json 2> gpurun_out/exp7_l1.err
FB200_LANES=0 python bench.py --no-cpu-baseline --steps 50 --warmup 10 > gpurun_out/exp7_l0.json 2> gpurun_out/exp7_l0.err
FB200_LANES=1 python bench.py --no-cpu-baseline --workload cfg5 --steps 50 --warmup 10 > gpurun_out/exp7_cfg5_l1.json 2> gpurun_out/exp7_cfg5.err
FB200_LANES=0 python bench.py --no-cpu-baseline --workload cfg5 --steps 50 --warmup 10 > gpurun_out/exp7_cfg5_l0.json 2>> gpurun_out/exp7_cfg5.err
echo done
