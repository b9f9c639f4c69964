#!/bin/bash
# Rebuild the library with FT kernel variants (-D switches of csrc/u1_ft.cu) and time one FT force.
export U1FT_CHUNK_SITES=4000000
for cfg in "-DU1_EPI_PREFETCH=0 -DU1_EPI_BATCH=0 -DU1_TAB_CPASYNC=0" "-DU1_EPI_PREFETCH=1 -DU1_EPI_BATCH=0 -DU1_TAB_CPASYNC=0" "-DU1_EPI_PREFETCH=0 -DU1_EPI_BATCH=1 -DU1_TAB_CPASYNC=0" "-DU1_EPI_PREFETCH=0 -DU1_EPI_BATCH=0 -DU1_TAB_CPASYNC=1" "-DU1_EPI_PREFETCH=1 -DU1_EPI_BATCH=1 -DU1_TAB_CPASYNC=0" "-DU1_EPI_PREFETCH=1 -DU1_EPI_BATCH=0 -DU1_TAB_CPASYNC=1"; do
  U1_NVCC_DEFINES="$cfg" python -m hmc_ft_b200.build --force > /dev/null 2>&1
  echo -n "[$cfg] : "; python tools/prof_ft.py rnet3 342 3
done
python -m hmc_ft_b200.build --force > /dev/null 2>&1
