#!/bin/bash
# Runs on the GPU box (under gpurun): full ncu capture of one persistent launch of the ring-mode-2 kernel on an HBM-bound
# narrow shape (default 4M x 32), plus the same shape on the shared-memory ring for comparison.
# Usage: tools/gpu_profile_ldg.sh <tag> [N] [D]
set -u
TAG=$1; N=${2:-4000000}; D=${3:-32}
mkdir -p gpurun_out
for rm in 2 1; do
  EDHMC_RING=$rm ncu --set full --clock-control none --import-source on -k regex:k_hmc -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_ring${rm} \
    python tools/quick_bench.py --N $N --D $D --T 3 --L 10 --reps 1 > gpurun_out/prof_${TAG}_ring${rm}.log 2>&1
  ncu -i gpurun_out/prof_${TAG}_ring${rm}.ncu-rep --page raw --csv > gpurun_out/${TAG}_ring${rm}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_ring${rm}.ncu-rep --page details > gpurun_out/${TAG}_ring${rm}_details.txt 2>/dev/null
done
ls -la gpurun_out/ | grep ${TAG}
