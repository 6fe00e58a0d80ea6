#!/bin/bash
# r02d: cluster split-K of the tcgen05 GEMM - parity first, then the batch sweep with it on / off
set -u
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_primitives.py -q -k "tcgen05" 2>&1 | tail -4 > $O/r02d_csplit_prim.log
cat $O/r02d_csplit_prim.log
timeout 200 python tools/r02d_gemm_sweep.py 2>&1 | grep -v "ok 0" | tail -8
timeout 900 python -m pytest tests/test_gpu_parity.py -q 2>&1 | tail -4 > $O/r02d_csplit_parity.log
cat $O/r02d_csplit_parity.log
for c in 8 0; do
  FB200_TC_CSPLIT=$c timeout 300 python bench.py --no-cpu-baseline --no-incumbent --no-extras --sweep 64,128,256,512,1024 --steps 30 > $O/r02d_csplit${c}_cfg2.json 2>$O/r02d_csplit${c}_cfg2.err
  FB200_TC_CSPLIT=$c timeout 300 python bench.py --workload cfg5 --no-cpu-baseline --no-incumbent --no-extras --sweep 32,128,256,1024 --steps 30 > $O/r02d_csplit${c}_cfg5.json 2>$O/r02d_csplit${c}_cfg5.err
done
python tools/show_bench.py $O/r02d_csplit8_cfg2.json $O/r02d_csplit0_cfg2.json $O/r02d_csplit8_cfg5.json $O/r02d_csplit0_cfg5.json 2>&1 | grep -v incumbent
