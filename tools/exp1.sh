#!/usr/bin/env bash
# experiment: 8 worker warps + PDL on/off
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/exp1_pytest.log
FB200_PDL=0 python bench.py --no-cpu-baseline --steps 50 --warmup 10 > gpurun_out/exp1_pdl0.json 2> gpurun_out/exp1_pdl0.err
FB200_PDL=1 python bench.py --no-cpu-baseline --steps 50 --warmup 10 > gpurun_out/exp1_pdl1.json 2> gpurun_out/exp1_pdl1.err
python tools/tc_trace.py 1 0 4096 512 2048 2>&1 | tail -14 > gpurun_out/exp1_trace.log
python tools/tc_trace.py 1 0 4096 512 512 2>&1 | tail -5 >> gpurun_out/exp1_trace.log
echo done
