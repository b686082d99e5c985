#!/bin/bash
# ncu launch list + full capture of the two wide-model GEMM kernels (run under gpurun, one GPU).
# usage: tools/gpu_profile_wide.sh <tag> [N] [D] [C]
TAG=${1:-r01f_wide}; N=${2:-200000}; D=${3:-1000}; C=${4:-1024}
mkdir -p gpurun_out
CMD="python tools/chain_bench.py --N $N --D $D --C $C --T 1 --L 2 --reps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launch_run.log 2>&1
for k in 1 2; do
  # launches alternate gemm<1>, gemm<2>; skip the initial evaluation and the first step
  ncu --set full --clock-control none --import-source on -k regex:k_mcw_gemm -s $((3 + k)) -c 1 -f -o gpurun_out/${TAG}_gemm${k} $CMD > gpurun_out/${TAG}_gemm${k}_run.log 2>&1
  ncu -i gpurun_out/${TAG}_gemm${k}.ncu-rep --page details > gpurun_out/${TAG}_gemm${k}_details.txt 2>&1
  ncu -i gpurun_out/${TAG}_gemm${k}.ncu-rep --page raw --csv > gpurun_out/${TAG}_gemm${k}_raw.csv 2>&1
done
