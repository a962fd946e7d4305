#!/bin/bash
# ncu launch list + full capture of the persistent MPC kernel (and optionally the forecast GEMM) for the current build.
# Usage (under gpurun, repo root): bash scripts/gpu_profile.sh <tag> [gemm]
TAG=${1:-r2}
PROF="python bench.py --paths 592 --steps 1 --warmup 3 --no-cpu-baseline"
$PROF > gpurun_out/${TAG}_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $PROF > gpurun_out/${TAG}_ncu1.log 2>&1
$PROF > gpurun_out/${TAG}_prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:backtest_lane -s 3 -c 1 -o gpurun_out/${TAG}_backtest_lane $PROF > gpurun_out/${TAG}_ncu2.log 2>&1
if [ "$2" = "gemm" ]; then
$PROF > gpurun_out/${TAG}_prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc16 -s 40 -c 4 -o gpurun_out/${TAG}_gemm_tc16 $PROF > gpurun_out/${TAG}_ncu3.log 2>&1
fi
echo "profile done rc=$?"
