#!/bin/bash
# Rebuild the library with FT kernel variants (-D switches of csrc/u1_ft.cu) and time one FT force.
for cfg in "-DU1_ADD_AT_START=1" "-DU1_ADD_AT_START=2"; do
  U1_NVCC_DEFINES="$cfg" python -m hmc_ft_b200.build --force > /dev/null 2>&1
  echo -n "[$cfg] : "; python tools/prof_ft.py rnet3 1024 2
done
python -m hmc_ft_b200.build --force > /dev/null 2>&1
