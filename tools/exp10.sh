#!/usr/bin/env bash
cd "$(dirname "$0")/.."
: > gpurun_out/exp10.log
for dbg in 119 0; do
  echo "== FB200_TC_DBG=$dbg" >> gpurun_out/exp10.log
  FB200_TC_DBG=$dbg python tools/tc_trace.py 1 0 4096 512 2048 2>&1 | tail -6 >> gpurun_out/exp10.log
done
