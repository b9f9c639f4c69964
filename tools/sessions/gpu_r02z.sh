#!/bin/bash
# ncu --set full of the final sparse final-layer kernels and the 8-rows 12->6 reverse layer (342 chains, one force).
tag=${1:-r02_z}
o=gpurun_out
U1FT_GRAPH=0 python tools/prof_ft_once.py 342 1 > $o/${tag}_ft_once.log 2>&1 &&
U1FT_GRAPH=0 ncu --set full --clock-control none --import-source on -k 'regex:k_final_bwd_fused|k_final_update_tile' -s 14 -c 4 -o $o/${tag}_final python tools/prof_ft_once.py 342 1 > $o/${tag}_ncu_final.log 2>&1
U1FT_GRAPH=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_conv_tile<.int.64, .int.8, .int.12, .int.6, .int.2>' -s 1 -c 1 -o $o/${tag}_rows8 python tools/prof_ft_once.py 342 1 > $o/${tag}_ncu_rows8.log 2>&1
tail -3 $o/${tag}_ncu_final.log $o/${tag}_ncu_rows8.log; ls -la $o | grep ${tag}
