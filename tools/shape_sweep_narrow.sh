#!/bin/bash
for shape in "16000000 8" "8000000 16" "6000000 24" "4000000 32" "4000000 40" "3000000 48" "2000000 54" "3000000 64" "3000000 72"; do
  set -- $shape
  for bt in 0 1; do
    echo -n "N=$1 D=$2 bigtile=$bt: "
    EDHMC_BIGTILE=$bt timeout 120 python tools/quick_bench.py --N $1 --D $2 --T 3 --L 10 --reps 3 | grep -E "steps/s" | sed 's/.*L=10: //' | cut -c1-110
  done
done
