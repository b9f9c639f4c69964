#!/bin/bash
# Rebuild the library with the FT kernel variants (GELU table version, forward tile rows) and time one FT force.
export U1FT_CHUNK_SITES=4000000
for cfg in "-DU1_GELU_V=1" "-DU1_GELU_V=2" "-DU1_GELU_V=2 -DU1_TILE_FWD_ROWS=4 -DU1_TILE_FWD_MINB=4" "-DU1_GELU_V=2 -DU1_TILE_FWD_ROWS=4 -DU1_TILE_FWD_MINB=3" "-DU1_GELU_V=2 -DU1_TILE_FWD_MINB=5"; do
  U1_NVCC_DEFINES="$cfg" python -m hmc_ft_b200.build --force > /dev/null 2>&1
  echo -n "[$cfg] : "; python tools/prof_ft.py rnet3 342 3
done
python -m hmc_ft_b200.build --force > /dev/null 2>&1
