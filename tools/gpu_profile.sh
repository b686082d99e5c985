#!/bin/bash
# Runs on the GPU box (under gpurun): launch list + one full ncu capture of the dominant kernel.
# Usage: tools/gpu_profile.sh <tag> [bench args...]
set -u
TAG=$1; shift
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline $*"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_hmc -s 3 -c 1 -f -o gpurun_out/prof_${TAG} $B > gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out/
