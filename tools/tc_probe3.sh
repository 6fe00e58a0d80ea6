#!/usr/bin/env bash
cd "$(dirname "$0")/.."
out=gpurun_out/tc_probe3.log; : > $out
for c in "1 0 128 128 64" "1 0 256 256 512" "1 0 4096 512 2048" "1 0 200 136 72" "1 0 40 512 512" "1 1 256 256 512" "1 1 4096 512 2048" "1 1 200 136 72" "1 1 257 512 256" "1 2 512 512 4096" "2 0 256 256 512"; do
  timeout 60 python tools/tc_probe.py $c >> $out 2>&1; echo "  -> exit $? ($c)" >> $out
done
grep -E "rel err|RC|rror|exit [^0]" $out
python tools/tc_trace.py 1 0 4096 512 2048 | tail -12
