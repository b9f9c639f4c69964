#!/bin/bash
# One-GPU checkpoint: all GPU tests, slab bench + profile, full bench line.
tag=${1:-r02_n}
o=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > $o/${tag}_tests.log
python bench.py --slab-only > $o/${tag}_slab1.json 2> $o/${tag}_slab1.err
python tools/slab_profile.py 4096 > $o/${tag}_slab_profile.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err
tail -n 5 $o/${tag}_tests.log
grep -o '"ms_per_step": [0-9.]*\|"parity": {[^}]*}' $o/${tag}_slab1.json | head -3
head -16 $o/${tag}_slab_profile.log | cut -c1-60,130-200
tail -2 $o/${tag}_slab_profile.log; tail -3 $o/${tag}_smoke.log; tail -c 300 $o/${tag}_bench.err
