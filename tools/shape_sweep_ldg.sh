#!/bin/bash
# Same-box A/B of ring mode 2 (bind-time re-lay + LDG, stream_ldg.cuh) against the shared-memory rings on narrow rows,
# at sizes that do not fit the L2 (HBM-bound) and at L2-resident sizes.
SHAPES=${SHAPES:-"16000000 8;8000000 16;6000000 24;4000000 32;4000000 33;4000000 40;3000000 48;2000000 54;3000000 64;581012 54;290506 54;1000000 8;500000 16;500000 32;400000 64"}
IFS=';' read -ra LIST <<< "$SHAPES"
for shape in "${LIST[@]}"; do
  set -- $shape
  for rm in 1 2; do
    echo -n "N=$1 D=$2 ring=$rm: "
    EDHMC_RING=$rm timeout 120 python tools/quick_bench.py --N $1 --D $2 --T 3 --L 10 --reps 3 | grep -E "steps/s" | sed 's/.*L=10: //' | cut -c1-110
  done
done
