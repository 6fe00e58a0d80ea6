#!/bin/bash
# Round-2 (fifth session), second evidence call after the classifier-head backward changed: the GPU suite, smoke(), the default
# bench line, the other BASELINE workloads, the launch list of the headline step.  (Timelines / ncu --set full: tools/final_r02e.sh.)
set -u
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r02e_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02e_smoke.log 2>&1
python bench.py > $O/r02e_bench_head.json 2> $O/r02e_bench_head.err
for wl in cfg3a cfg4b cfg5; do
  python bench.py --workload $wl --no-cpu-baseline --no-incumbent --no-extras --sweep "32,128,256,1024" --steps 30 > $O/r02e_bench_$wl.json 2> /dev/null
done
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-incumbent --no-extras --sweep="
$CMD > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_r02e.csv $CMD > $O/r02e_ncu_launches.log 2>&1
tail -3 $O/r02e_gputest.log; tail -2 $O/r02e_smoke.log; python tools/show_bench.py $O/r02e_bench_head.json 2>/dev/null | head -14
for wl in cfg3a cfg4b cfg5; do python tools/show_bench.py $O/r02e_bench_$wl.json 2>/dev/null | head -3; done
