#!/bin/bash
# r02d: validate the TabTransformer backward with shared memory under 164 KB (L1 keeps the layer weights) + packed-parameter autograd
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_tabt.py -q -x 2>&1 | tail -5 > $O/r02d_tabt_test.log
timeout 300 python tools/tabt_bench.py 32 1024 4096 > $O/r02d_tabt_bench.log 2>&1
timeout 120 python tools/tabt_trace.py > $O/r02d_tabt_trace.log 2>&1
cat $O/r02d_tabt_test.log; cat $O/r02d_tabt_bench.log | tail -4; tail -32 $O/r02d_tabt_trace.log
