#!/bin/bash
# tools/pair_ab.sh -- A/B of the single-step kernel against k_step_pair variants inside ONE gpurun call.
cd "$(dirname "$0")/.."
L=cfd-codes_b200/liblbm_b200.so
GRIDS=${GRIDS:-"256 512"}
VARS=${VARS:-"0 1 2 3 4 5"}
LZS=${LZS:-"16 32 64"}
for N in $GRIDS; do
  python tools/time_lib.py $L $N pair=0
  for v in $VARS; do
    for lz in $LZS; do
      python tools/time_lib.py $L $N pair=1 pair_variant=$v pair_lz=$lz
    done
  done
done
