set -u
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-incumbent --no-extras --steps 30 --batch 256"
$B --sweep 48,64,128,512,1024 > $O/r02d_exp_base.json 2>/dev/null
FB200_TAIL_FUSE=1 $B --sweep 48,64,128,512,1024 > $O/r02d_exp_tail.json 2>/dev/null
FB200_MEGA=0 $B --sweep 33,48,64 > $O/r02d_exp_nomega.json 2>/dev/null
FB200_MEGA=0 FB200_TAIL_FUSE=1 $B --sweep 33,48,64 > $O/r02d_exp_nomega_tail.json 2>/dev/null
python tools/show_bench.py $O/r02d_exp_base.json $O/r02d_exp_tail.json $O/r02d_exp_nomega.json $O/r02d_exp_nomega_tail.json | grep -v "incumbent\|clocks\|roofline:"
