#!/bin/bash
# A/B of the planner's big-tile rule over row widths (run under gpurun, one GPU)
for shape in "4000000 64" "2000000 128" "1500000 200" "1000000 300" "800000 512" "600000 777" "1250000 1000" "500000 1500" "400000 2048" "1000000 131" "3000000 96"; do
  set -- $shape
  for bt in 0 1; do
    echo -n "N=$1 D=$2 bigtile=$bt: "
    EDHMC_BIGTILE=$bt timeout 120 python tools/quick_bench.py --N $1 --D $2 --T 3 --L 10 --reps 3 | grep -E "steps/s" | sed 's/.*L=10: //' | cut -c1-110
  done
done
