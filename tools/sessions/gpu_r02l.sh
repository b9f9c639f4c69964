#!/bin/bash
# One-GPU session: slab trajectory with carried sums / fused reductions: unit tests, slab tests, bench, profile.
tag=${1:-r02_l}
o=gpurun_out
python -m pytest tests/test_plain_gpu.py -q -x -k "slab or multistep or rect" 2>&1 | tail -15 > $o/${tag}_tests_plain.log
python -m pytest tests/test_multi_gpu.py -q 2>&1 | tail -30 > $o/${tag}_tests_multi.log
python bench.py --slab-only > $o/${tag}_slab1.json 2> $o/${tag}_slab1.err
python tools/slab_profile.py 4096 > $o/${tag}_slab_profile.log 2>&1
tail -n 5 $o/${tag}_tests_plain.log $o/${tag}_tests_multi.log
grep -o '"ms_per_step": [0-9.]*\|"parity": {[^}]*}' $o/${tag}_slab1.json | head -3
head -24 $o/${tag}_slab_profile.log | cut -c1-60,130-200
tail -2 $o/${tag}_slab_profile.log
