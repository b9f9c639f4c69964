#!/bin/bash
# FT force timing for several chunk sizes (sites per launch); run on the GPU box.
for cs in 227328 681984 1818624; do
  B=$((cs / 4096))
  echo "chunk_sites=$cs B=$B"
  U1FT_CHUNK_SITES=$cs python tools/prof_ft.py rnet3 $B 3
done
U1FT_CHUNK_SITES=524288 python tools/prof_ft.py rnet3 128 3
U1FT_CHUNK_SITES=681984 python tools/prof_ft.py simple 166 3
U1FT_CHUNK_SITES=681984 python tools/prof_ft.py rnet1 166 3
