#!/bin/bash
# One-GPU session: all GPU tests + smoke on the final code, slab bench + profile (exchange kernel with 4 loads in flight,
# merged boundary launch), ncu evidence of the final FT kernels (launch list of one force, --set full of the forward conv with
# the v3 GELU table and of the two fused final-layer kernels).
tag=${1:-r02_s}
o=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > $o/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1
python bench.py --slab-only > $o/${tag}_slab1.json 2> $o/${tag}_slab1.err
python tools/slab_profile.py 4096 > $o/${tag}_slab_profile.log 2>&1
U1FT_GRAPH=0 python tools/prof_ft_once.py 342 1 > $o/${tag}_ft_once.log 2>&1 &&
U1FT_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/${tag}_launches_ft.csv python tools/prof_ft_once.py 342 1 > /dev/null 2>&1
U1FT_GRAPH=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_conv_tile<.int.64, .int.4, .int.12, .int.12, .int.0>' -s 2 -c 2 -o $o/${tag}_conv_fwd python tools/prof_ft_once.py 342 1 > $o/${tag}_ncu_fwd.log 2>&1
U1FT_GRAPH=0 ncu --set full --clock-control none --import-source on -k 'regex:k_final_update|k_final_bwd_fused' -s 2 -c 2 -o $o/${tag}_final python tools/prof_ft_once.py 342 1 > $o/${tag}_ncu_final.log 2>&1
tail -n 4 $o/${tag}_tests.log; tail -2 $o/${tag}_smoke.log
grep -o '"ms_per_step": [0-9.]*\|"parity": {[^}]*}' $o/${tag}_slab1.json | head -3
head -16 $o/${tag}_slab_profile.log | cut -c1-60,130-200; tail -2 $o/${tag}_slab_profile.log
ls -la $o | grep ${tag}
