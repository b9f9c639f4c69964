#!/bin/bash
# FT force timing for several batch sizes (single chunk each); run on the GPU box.
export U1FT_CHUNK_SITES=4000000
for B in 16 64 128 342 444; do
  python tools/prof_ft.py rnet3 $B 3
done
python tools/prof_ft.py simple 128 3
python tools/prof_ft.py rnet1 128 3
