#!/usr/bin/env bash
# where does the fp32-strict mainloop time go?  timing-only variants (results wrong by construction)
cd "$(dirname "$0")/.."
: > gpurun_out/exp8.log
for dbg in 0 1 2 3 4 8 7; do
  echo "== FB200_TC_DBG=$dbg" >> gpurun_out/exp8.log
  FB200_TC_DBG=$dbg python tools/tc_trace.py 1 0 4096 512 2048 2>&1 | tail -4 >> gpurun_out/exp8.log
done
echo "== layout 2 (dW) dbg 0" >> gpurun_out/exp8.log
python tools/tc_trace.py 1 2 512 2048 4096 2>&1 | tail -4 >> gpurun_out/exp8.log
echo "== bf16 engine" >> gpurun_out/exp8.log
python tools/tc_trace.py 2 0 4096 512 2048 2>&1 | tail -4 >> gpurun_out/exp8.log
