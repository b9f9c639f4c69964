#!/bin/bash
# Cluster / DSMEM form of the multistep kernel: bit-identity + timing at cluster sizes 0 (off), 2, 4, 8.
tag=${1:-r02_q}
o=gpurun_out
for cl in 0 2 4 8; do
  echo "U1_MULTISTEP_CL=$cl" >> $o/${tag}_cl.log
  U1_MULTISTEP_CL=$cl timeout 300 python -m pytest tests/test_plain_gpu.py -q -k "multistep or rect or slab" 2>&1 | tail -3 >> $o/${tag}_cl.log
  for lx in 4096 560 2096; do U1_MULTISTEP_CL=$cl timeout 120 python tools/prof_multi.py 4096 $lx; done >> $o/${tag}_cl.log 2>&1
done
cat $o/${tag}_cl.log
