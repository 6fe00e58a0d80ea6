#!/usr/bin/env bash
# runs the tcgen05 probe matrix, one process per case, each under its own timeout
cd "$(dirname "$0")/.."
out=gpurun_out/tc_probe.log; mkdir -p gpurun_out; : > $out
for eng in ${ENGINES:-1 2}; do for lay in 0 1 2; do
  for shape in "128 128 64" "256 256 512" "4096 512 2048" "512 512 4096" "200 136 72" "40 512 512" "4096 2048 512"; do
    timeout 60 python tools/tc_probe.py $eng $lay $shape >> $out 2>&1; echo "  -> exit $? ($eng $lay $shape)" >> $out
  done
  timeout 60 python tools/tc_probe.py $eng $lay 256 256 512 nobias >> $out 2>&1; echo "  -> exit $? ($eng $lay nobias)" >> $out
done; done
grep -E "rel err|exit|RC|rror" $out | tail -120
