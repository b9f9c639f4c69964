#!/bin/bash
# One-GPU session: FT fusions (final conv + update, features in the first conv) and GELU table v3: parity tests, A/B.
tag=${1:-r02_j}
o=gpurun_out
python -m pytest tests/test_ft_gpu.py tests/test_train_gpu.py -q -x 2>&1 | tail -15 > $o/${tag}_tests_ft.log
U1FT_FUSE=0 U1FT_FUSE_FEAT=0 python -m pytest tests/test_ft_gpu.py -q -x 2>&1 | tail -5 > $o/${tag}_tests_ft_nofuse.log
for cfg in "1 1" "0 1" "1 0" "0 0" "1 1"; do set -- $cfg
  echo "FUSE=$1 FEAT=$2" >> $o/${tag}_force.log
  U1FT_FUSE=$1 U1FT_FUSE_FEAT=$2 python tools/prof_ft.py rnet3 342 10 >> $o/${tag}_force.log 2>&1
  U1FT_FUSE=$1 U1FT_FUSE_FEAT=$2 python tools/prof_ft.py rnet3 1024 5 >> $o/${tag}_force.log 2>&1
done
for lanes in 1 2; do for B in 1024 512; do echo "LANES=$lanes" >> $o/${tag}_lanes.log; U1FT_LANES=$lanes python tools/prof_ft_traj.py $B 10 >> $o/${tag}_lanes.log 2>&1; done; done
python tools/ft_small_batch.py 128 > $o/${tag}_ft128.log 2>&1
tail -4 $o/${tag}_tests_ft.log $o/${tag}_tests_ft_nofuse.log; cat $o/${tag}_force.log $o/${tag}_lanes.log $o/${tag}_ft128.log
