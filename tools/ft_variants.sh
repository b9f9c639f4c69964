#!/bin/bash
# Rebuild the library once per configuration (-D switches of csrc/u1_ft.cu, e.g. "-DU1_GELU_V=1",
# "-DU1_TILE_FWD_ROWS=4 -DU1_TILE_FWD_MINB=4", "-DU1_ADD_AT_START=0", "-DU1_TAB_CPASYNC=0") and time one FT force.
#   tools/ft_variants.sh "<defines A>" "<defines B>" ...      (run on the GPU box)
for cfg in "$@"; do
  U1_NVCC_DEFINES="$cfg" python -m hmc_ft_b200.build --force > /dev/null 2>&1
  echo -n "[$cfg] : "; python tools/prof_ft.py rnet3 1024 2
done
python -m hmc_ft_b200.build --force > /dev/null 2>&1
