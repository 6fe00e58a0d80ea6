set -x
python -m pytest tests/test_gpu_tabt.py -q -k deterministic 2>&1 | tail -3
python tools/attn_once.py && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02c_attn_launches.csv python tools/attn_once.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 3 -c 3 -o gpurun_out/r02c_attn_tc -f python tools/attn_once.py > gpurun_out/r02c_attn_tc_ncu.log 2>&1
python tools/tabt_once.py 1024 && ncu --set full --clock-control none --import-source on -k regex:tabt_ -s 3 -c 3 -o gpurun_out/r02c_tabt -f python tools/tabt_once.py 1024 > gpurun_out/r02c_tabt_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
