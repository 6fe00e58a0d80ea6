#!/bin/bash
# Round-2 (fifth session) evidence, one gpurun call on one B200: the whole GPU suite, smoke(), the default bench line, the other
# BASELINE workloads with their batch sweeps, the grid-wide GEMM timelines, the launch lists and the ncu capture of the dominant
# kernel (every profiled command first exits 0 without the profiler).  `python tools/summarize_profiles.py r02e` -> profiles/r02e_*.
set -u
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r02e_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02e_smoke.log 2>&1
python bench.py > $O/r02e_bench_head.json 2> $O/r02e_bench_head.err
for wl in cfg3a cfg4b cfg5; do
  python bench.py --workload $wl --no-cpu-baseline --no-incumbent --no-extras --sweep "32,128,256,1024" --steps 30 > $O/r02e_bench_$wl.json 2> /dev/null
done
python tools/tc_timeline.py cfg2 4096 > $O/r02e_timeline_cfg2_B4096.txt 2>&1
python tools/tc_timeline.py cfg5 4096 > $O/r02e_timeline_cfg5_B4096.txt 2>&1
python tools/tc_timeline.py cfg2 256 > $O/r02e_timeline_cfg2_B256.txt 2>&1
( python tools/tc_handover.py 1 0 16384 512 512; python tools/tc_handover.py 1 1 16384 512 512; python tools/tc_handover.py 2 0 16384 512 512 ) > $O/r02e_handover_gemm.txt 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-incumbent --no-extras --sweep="
$CMD > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_r02e.csv $CMD > $O/r02e_ncu_launches.log 2>&1
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 45 -c 6 -o $O/prof_r02e_tc -f $CMD > $O/r02e_ncu_full.log 2>&1
CMD5="python bench.py --workload cfg5 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-incumbent --no-extras --sweep="
$CMD5 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_r02e_cfg5.csv $CMD5 > $O/r02e_ncu_cfg5.log 2>&1
tail -3 $O/r02e_gputest.log; tail -2 $O/r02e_smoke.log; python tools/show_bench.py $O/r02e_bench_head.json 2>/dev/null | head -14
for wl in cfg3a cfg4b cfg5; do python tools/show_bench.py $O/r02e_bench_$wl.json 2>/dev/null | head -5; done
tail -4 $O/r02e_handover_gemm.txt; ls -la $O/*.ncu-rep | tail -2
