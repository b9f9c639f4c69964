#!/bin/bash
# One-GPU session of round 2: all GPU tests, slab-only bench, full bench, then the ncu evidence of the SHIPPED kernels
# (launch lists + --set full of the forward/reverse tiled conv, the resident trajectory kernel, the multistep kernel).
tag=${1:-r02_h}
o=gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > $o/${tag}_tests.log
python bench.py --slab-only > $o/${tag}_slab1.json 2> $o/${tag}_slab1.err
python bench.py --steps 5 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err
M=gpu__time_duration.sum
U1FT_GRAPH=0 python tools/prof_ft_once.py 342 1 > $o/${tag}_ft_once.log 2>&1 &&
U1FT_GRAPH=0 ncu --metrics $M --clock-control none --csv --log-file $o/${tag}_launches_ft.csv python tools/prof_ft_once.py 342 1 > /dev/null 2>&1
U1FT_GRAPH=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_conv_tile<.int.64, .int.4, .int.12, .int.12, .int.0>" -s 2 -c 2 -o $o/${tag}_conv_fwd python tools/prof_ft_once.py 342 1 > $o/${tag}_ncu_fwd.log 2>&1
U1FT_GRAPH=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_conv_tile<.int.64, .int.4, .int.12, .int.12, .int.2>" -s 2 -c 2 -o $o/${tag}_conv_bwd python tools/prof_ft_once.py 342 1 > $o/${tag}_ncu_bwd.log 2>&1
python tools/prof_counters.py > $o/${tag}_counters.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_traj_resident" -s 2 -c 1 -o $o/${tag}_resident python tools/prof_counters.py > $o/${tag}_ncu_res.log 2>&1
python tools/prof_multi.py 4096 560 > $o/${tag}_multi.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_leapfrog_multi" -s 10 -c 1 -o $o/${tag}_multi python tools/prof_multi.py 4096 560 > $o/${tag}_ncu_multi.log 2>&1
tail -5 $o/${tag}_tests.log
grep -o '"ms_per_step": [0-9.]*' $o/${tag}_slab1.json | head -2
tail -c 600 $o/${tag}_bench.err
cat $o/${tag}_multi.log
ls -la $o | grep ${tag}
