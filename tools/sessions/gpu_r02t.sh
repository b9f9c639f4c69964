#!/bin/bash
# FT final-layer kernels: tiled k_final_update, two-pass k_final_bwd_fused: parity + timing + launch list.
tag=${1:-r02_t}
o=gpurun_out
python -m pytest tests/test_ft_gpu.py tests/test_train_gpu.py -q -x 2>&1 | tail -8 > $o/${tag}_tests_ft.log
for i in 1 2; do python tools/prof_ft.py rnet3 342 10; python tools/prof_ft.py rnet3 1024 5; done > $o/${tag}_force.log 2>&1
python tools/prof_ft.py simple 1024 10 >> $o/${tag}_force.log 2>&1
U1FT_GRAPH=0 python tools/prof_ft_once.py 342 1 > $o/${tag}_ft_once.log 2>&1 &&
U1FT_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/${tag}_launches_ft.csv python tools/prof_ft_once.py 342 1 > /dev/null 2>&1
tail -n 3 $o/${tag}_tests_ft.log; cat $o/${tag}_force.log
