#!/bin/bash
# r02d step 2: class-head backward split (dX kernel + deferred dW GEMM), bf16 plain epilogue, step kernel up to 32 rows
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_primitives.py tests/test_gpu_pdl.py -q 2>&1 | tail -4 > $O/r02d_step2_tests.log
cat $O/r02d_step2_tests.log
B="python bench.py --no-cpu-baseline --no-incumbent --no-extras --steps 30"
$B --sweep 32,64,256,1024 > $O/r02d_s2_cfg2.json 2>/dev/null
FB200_HEAD_SPLIT=0 $B --sweep 256,1024 > $O/r02d_s2_cfg2_nosplit.json 2>/dev/null
$B --workload cfg5 --sweep 32,128,1024 > $O/r02d_s2_cfg5.json 2>/dev/null
$B --workload cfg4b --sweep 32,128,1024 > $O/r02d_s2_cfg4b.json 2>/dev/null
$B --workload cfg3a --sweep 32,256,1024 > $O/r02d_s2_cfg3a.json 2>/dev/null
python tools/show_bench.py $O/r02d_s2_cfg2.json $O/r02d_s2_cfg2_nosplit.json $O/r02d_s2_cfg5.json $O/r02d_s2_cfg4b.json $O/r02d_s2_cfg3a.json | grep -v "incumbent\|clocks"
