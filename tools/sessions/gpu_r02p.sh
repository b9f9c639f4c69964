#!/bin/bash
# FT small-batch tuning: rows-per-thread threshold together with lanes (per-lane launches are half the size).
tag=${1:-r02_p}
o=gpurun_out
for cfg in "600 2" "500 2" "200 4" "200 2" "600 2"; do set -- $cfg
  echo "U1FT_TILE_T4=$1 U1FT_LANES=$2" >> $o/${tag}_rows_lanes.log
  U1FT_TILE_T4=$1 U1FT_LANES=$2 python tools/ft_small_batch.py 128 >> $o/${tag}_rows_lanes.log 2>&1
done
echo "256 chains: T4=600 auto lanes, then T4=500" >> $o/${tag}_rows_lanes.log
python tools/ft_small_batch.py 256 >> $o/${tag}_rows_lanes.log 2>&1
U1FT_TILE_T4=500 python tools/ft_small_batch.py 256 >> $o/${tag}_rows_lanes.log 2>&1
cat $o/${tag}_rows_lanes.log
