#!/bin/bash
# ncu evidence for the many-chain tensor-core pass (under gpurun). Usage: tools/gpu_profile_chains.sh <tag>
TAG=$1
mkdir -p gpurun_out
B="python tools/chain_bench.py --T 1 --L 4 --reps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mc_pass_tc3 -s 2 -c 1 -f -o gpurun_out/prof_${TAG} $B > gpurun_out/prof_${TAG}.log 2>&1
