#!/bin/bash
# FT force timing for several chunk sizes (sites per launch); run on the GPU box.
for cs in 909312 1818624 3637248; do
  B=$((cs / 4096))
  echo "chunk_sites=$cs B=$B"
  U1FT_CHUNK_SITES=$cs python tools/prof_ft.py rnet3 $B 3
done
