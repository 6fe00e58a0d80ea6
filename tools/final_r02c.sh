#!/bin/bash
# Round-2 (third session) evidence, one gpurun call on one B200: the whole GPU suite, the default bench line (with the
# TabTransformer / token-attention extras), micro-benchmarks of the two f3 kernels against stock torch.nn on the same GPU,
# and their ncu captures (every profiled command first exits 0 without the profiler).
# Outputs under gpurun_out/; `python tools/summarize_profiles.py r02c` turns them into profiles/r02c_*.
set -u
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r02c_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02c_smoke.log 2>&1
python bench.py > $O/r02c_bench_head.json 2> $O/r02c_bench_head.err
python tools/tabt_bench.py 32 1024 4096 > $O/r02c_tabt_bench.log 2>&1
(python tools/attn_bench.py 32; FB200_ATTN_TC=0 python tools/attn_bench.py 32; python tools/attn_bench.py 256; FB200_ATTN_TC=0 python tools/attn_bench.py 256) 2>/dev/null > $O/r02c_attn_bench.log
python tools/attn_once.py > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 3 -c 3 -o $O/r02c_attn_tc -f python tools/attn_once.py > $O/r02c_attn_tc_ncu.log 2>&1
python tools/tabt_once.py 4096 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tabt_ -s 3 -c 3 -o $O/r02c_tabt -f python tools/tabt_once.py 4096 > $O/r02c_tabt_ncu.log 2>&1
tail -3 $O/r02c_gputest.log; cat $O/r02c_smoke.log | tail -2; python tools/show_bench.py $O/r02c_bench_head.json 2>/dev/null | head -12
