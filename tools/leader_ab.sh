#!/bin/bash
# Same-box A/B of the serial-section protocol of the persistent plan on cfg 2: grid barrier (EDHMC_LEADER=0) vs leader.
for v in 0 1 0 1; do echo "EDHMC_LEADER=$v"; EDHMC_LEADER=$v timeout 120 python tools/quick_bench.py --reps 7 | tail -n 2; done
for n in 100000 290506 2000000; do for v in 0 1; do echo "N=$n EDHMC_LEADER=$v"; EDHMC_LEADER=$v timeout 120 python tools/quick_bench.py --N $n --reps 5 | tail -n 2 | head -n 1; done; done
