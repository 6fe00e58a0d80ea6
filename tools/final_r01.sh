#!/usr/bin/env bash
# round-1 final measurements: tests, bench (own + reference arm), ncu launch list and full captures
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/final_pytest.log
python bench.py --steps 50 --warmup 10 --sweep 32,256,1024,4736,16384 > $O/bench_r01_final.json 2> $O/bench_r01_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_r01_reference.json 2> $O/bench_r01_reference.err; echo "ref rc=$?"
python bench.py --no-cpu-baseline --workload cfg5 --steps 50 --warmup 10 > $O/bench_r01_cfg5.json 2> /dev/null
python bench.py --no-cpu-baseline --workload cfg3a --steps 50 --warmup 10 > $O/bench_r01_cfg3a.json 2> /dev/null
python bench.py --no-cpu-baseline --workload cfg4b --steps 50 --warmup 10 > $O/bench_r01_cfg4b.json 2> /dev/null
# launch list of one eager train step (bench exits 0 without ncu first: the runs above)
python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $O/plain_r01.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_list.log 2>&1; echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -c 6 -f -o $O/prof_r01_tc python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "full rc=$?"
python tools/attn_once.py > $O/attn_plain.log 2>&1; echo "attn rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_ -c 3 -f -o $O/prof_r01_attn python tools/attn_once.py > $O/ncu_attn.log 2>&1; echo "attn ncu rc=$?"
