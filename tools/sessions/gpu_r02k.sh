#!/bin/bash
# One-GPU session: reverse-sweep fusion parity + A/B, automatic lanes, slab per-kernel profile.
tag=${1:-r02_k}
o=gpurun_out
python -m pytest tests/test_ft_gpu.py tests/test_train_gpu.py -q -x 2>&1 | tail -15 > $o/${tag}_tests_ft.log
for cfg in 1 0 1 0; do
  echo "FUSE_BWD=$cfg" >> $o/${tag}_force.log
  U1FT_FUSE_BWD=$cfg python tools/prof_ft.py rnet3 342 10 >> $o/${tag}_force.log 2>&1
  U1FT_FUSE_BWD=$cfg python tools/prof_ft.py rnet3 1024 5 >> $o/${tag}_force.log 2>&1
done
python tools/ft_small_batch.py 128 > $o/${tag}_ft128.log 2>&1
python tools/slab_profile.py 4096 > $o/${tag}_slab_profile.log 2>&1
tail -n 4 $o/${tag}_tests_ft.log; cat $o/${tag}_force.log $o/${tag}_ft128.log; head -40 $o/${tag}_slab_profile.log | cut -c1-200
