#!/usr/bin/env bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/exp2_pytest.log
FB200_PDL=1 python bench.py --no-cpu-baseline --steps 50 --warmup 10 > gpurun_out/exp2_pdl1.json 2> gpurun_out/exp2_pdl1.err
FB200_PDL=0 python bench.py --no-cpu-baseline --steps 50 --warmup 10 > gpurun_out/exp2_pdl0.json 2> gpurun_out/exp2_pdl0.err
python bench.py --no-cpu-baseline --workload cfg5 --steps 50 --warmup 10 > gpurun_out/exp2_cfg5.json 2> gpurun_out/exp2_cfg5.err
python tools/tc_trace.py 1 0 4096 512 2048 2>&1 | tail -14 > gpurun_out/exp2_trace.log
python tools/tc_probe.py 1 2 512 512 4096 >> gpurun_out/exp2_trace.log 2>&1
python tools/tc_probe.py 1 1 4096 512 512 >> gpurun_out/exp2_trace.log 2>&1
echo done
