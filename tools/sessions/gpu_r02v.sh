#!/bin/bash
# 12->6 reverse layer with 8 rows per thread: parity + A/B.
tag=${1:-r02_v}
o=gpurun_out
python -m pytest tests/test_ft_gpu.py tests/test_train_gpu.py -q -x 2>&1 | tail -8 > $o/${tag}_tests_ft.log
for r8 in 1 0 1 0; do echo "U1FT_ROWS8=$r8" >> $o/${tag}_force.log
  U1FT_ROWS8=$r8 python tools/prof_ft.py rnet3 342 10 >> $o/${tag}_force.log 2>&1
  U1FT_ROWS8=$r8 python tools/prof_ft.py rnet3 1024 5 >> $o/${tag}_force.log 2>&1
  U1FT_ROWS8=$r8 python tools/prof_ft.py simple 1024 10 >> $o/${tag}_force.log 2>&1
done
tail -n 3 $o/${tag}_tests_ft.log; cat $o/${tag}_force.log
