#!/bin/bash
# Rebuild the library with different register caps for the tiled conv kernels and time one FT force.
export U1FT_CHUNK_SITES=4000000
for f in 5 6 7; do for b in 3 4; do
  U1_NVCC_DEFINES="-DU1_TILE_FWD_MINB=$f -DU1_TILE_BWD_MINB=$b" python -m hmc_ft_b200.build --force > /dev/null 2>&1
  echo -n "fwd_minb=$f bwd_minb=$b : "; python tools/prof_ft.py rnet3 342 3
done; done
python -m hmc_ft_b200.build --force > /dev/null 2>&1
