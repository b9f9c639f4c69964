#!/bin/bash
# One-GPU session: multistep-kernel prefetch variants (A/B + bit-identity), slab tests, FT lanes at small batch, conv ncu.
tag=${1:-r02_i}
o=gpurun_out
for pf in 0 1 2; do
  U1_MULTISTEP_PF=$pf python -m pytest tests/test_plain_gpu.py -q -k "multistep or rect or slab" 2>&1 | tail -3 > $o/${tag}_tests_pf$pf.log
  for lx in 4096 560 2096; do U1_MULTISTEP_PF=$pf python tools/prof_multi.py 4096 $lx; done > $o/${tag}_multi_pf$pf.log 2>&1
done
python -m pytest tests/test_multi_gpu.py -q 2>&1 | tail -30 > $o/${tag}_tests_multi.log
for pf in 0 1 2; do U1_MULTISTEP_PF=$pf python bench.py --slab-only > $o/${tag}_slab1_pf$pf.json 2> $o/${tag}_slab1_pf$pf.err; done
python tools/ft_small_batch.py 128 > $o/${tag}_ft128_lanes1.log 2>&1
U1FT_LANES=2 python tools/ft_small_batch.py 128 > $o/${tag}_ft128_lanes2.log 2>&1
U1FT_LANES=2 python tools/ft_small_batch.py 256 > $o/${tag}_ft256_lanes2.log 2>&1
U1FT_GRAPH=0 python tools/prof_ft_once.py 342 1 > $o/${tag}_ft_once.log 2>&1 &&
U1FT_GRAPH=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_conv_tile<.int.64, .int.4, .int.12, .int.12, .int.0>' -s 2 -c 2 -o $o/${tag}_conv_fwd python tools/prof_ft_once.py 342 1 > $o/${tag}_ncu_fwd.log 2>&1
U1FT_GRAPH=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_conv_tile<.int.64, .int.4, .int.12, .int.12, .int.2>' -s 2 -c 2 -o $o/${tag}_conv_bwd python tools/prof_ft_once.py 342 1 > $o/${tag}_ncu_bwd.log 2>&1
tail -3 $o/${tag}_tests_pf*.log $o/${tag}_tests_multi.log
cat $o/${tag}_multi_pf*.log
grep -o '"ms_per_step": [0-9.]*' $o/${tag}_slab1_pf*.json
cat $o/${tag}_ft128_lanes1.log $o/${tag}_ft128_lanes2.log $o/${tag}_ft256_lanes2.log
tail -3 $o/${tag}_ncu_fwd.log
