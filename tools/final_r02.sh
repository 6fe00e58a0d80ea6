#!/bin/bash
# Round-2 evidence, one gpurun call on one B200: GPU test suite, the default bench line, the reference arm, the other
# BASELINE workloads, and the ncu captures (each profiled command first exits 0 without the profiler).
# Outputs under gpurun_out/; tools/summarize_profiles.py r02 turns them into profiles/r02_*.
set -u
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r02_gputest.log
python bench.py > $O/r02_bench_head.json 2> $O/r02_bench_head.err
python bench.py --impl reference --steps 5 --warmup 2 > $O/r02_bench_reference.json 2> /dev/null
for wl in cfg3a cfg4b cfg5; do
  python bench.py --workload $wl --no-cpu-baseline --no-incumbent --no-extras --sweep "32,1024" --steps 30 > $O/r02_bench_$wl.json 2> /dev/null
done
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-incumbent --no-extras --sweep="
$CMD > $O/r02_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_r02.csv $CMD > $O/r02_ncu_launches.log 2>&1
$CMD > $O/r02_plain2.log 2>&1 &&
ncu --set full --clock-control none -k regex:tc_gemm -s 45 -c 6 -o $O/prof_r02_tc $CMD > $O/r02_ncu_full.log 2>&1
CMD32="python bench.py --batch 32 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-incumbent --no-extras --sweep="
$CMD32 > $O/r02_plain32.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:mega_step -c 4 --csv --log-file $O/launches_r02_b32.csv $CMD32 > $O/r02_ncu_b32.log 2>&1
ls -la $O/*.ncu-rep
du -sh $O
