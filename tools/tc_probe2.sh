#!/usr/bin/env bash
cd "$(dirname "$0")/.."
out=gpurun_out/tc_probe2.log; : > $out
for eng in 1 2; do
  for c in "1 257 512 256" "1 264 512 256" "2 256 512 257" "2 256 512 264" "2 256 512 33" "0 257 512 256" "2 256 512 129" "2 256 512 160" ; do
    timeout 60 python tools/tc_probe.py $eng $c nobias >> $out 2>&1; echo "  -> exit $? ($eng $c)" >> $out
  done
done
grep -E "rel err|RC|rror" $out
