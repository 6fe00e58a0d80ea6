#!/bin/bash
# Round-2 (fourth session) evidence, one gpurun call on one B200: the whole GPU suite, smoke(), the default bench line, the other
# BASELINE workloads with their batch sweeps, the cluster split-K A/B, launch lists and ncu captures of the changed kernels
# (every profiled command first exits 0 without the profiler).  `python tools/summarize_profiles.py r02d` -> profiles/r02d_*.
set -u
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r02d_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02d_smoke.log 2>&1
python bench.py > $O/r02d_bench_head.json 2> $O/r02d_bench_head.err
for wl in cfg3a cfg4b cfg5; do
  python bench.py --workload $wl --no-cpu-baseline --no-incumbent --no-extras --sweep "32,128,256,1024" --steps 30 > $O/r02d_bench_$wl.json 2> /dev/null
done
( for c in 8 0; do echo "FB200_TC_CSPLIT=$c"; FB200_TC_CSPLIT=$c python bench.py --no-cpu-baseline --no-incumbent --no-extras --sweep 64,128,256,512,1024 --steps 30 2>/dev/null | python tools/show_bench.py /dev/stdin | grep sweep; done ) > $O/r02d_csplit_ab.txt 2>&1
CMD="python bench.py --batch 256 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-incumbent --no-extras --sweep="
$CMD > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_r02d_b256.csv $CMD > $O/r02d_ncu_b256.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 30 -c 6 -o $O/r02d_tc_csplit -f $CMD > $O/r02d_ncu_csplit.log 2>&1
CMD5="python bench.py --workload cfg5 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-incumbent --no-extras --sweep="
$CMD5 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_r02d_cfg5.csv $CMD5 > $O/r02d_ncu_cfg5.log 2>&1
ncu --set full --clock-control none -k regex:grb_ -s 4 -c 4 -o $O/r02d_grb -f $CMD5 > $O/r02d_ncu_grb.log 2>&1
tail -3 $O/r02d_gputest.log; tail -2 $O/r02d_smoke.log; python tools/show_bench.py $O/r02d_bench_head.json 2>/dev/null | head -14; cat $O/r02d_csplit_ab.txt
ls -la $O/*.ncu-rep | tail -3
