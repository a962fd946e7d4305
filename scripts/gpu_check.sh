#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench line, then the ncu launch list + a full capture of the MPC kernel.
# Usage (from the repo root, under gpurun): bash scripts/gpu_check.sh [tag]
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
PROF="python bench.py --paths 1184 --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs"
$PROF > gpurun_out/${TAG}_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $PROF > gpurun_out/${TAG}_ncu1.log 2>&1
$PROF > gpurun_out/${TAG}_prof_plain2.log 2>&1 &&
# the five MPC launches of the timed step: dense start (backtest_lane_kernel), reduced solves (backtest_active_kernel), second chance (2), stragglers
ncu --set full --clock-control none --import-source on -k regex:backtest_ -s 15 -c 5 -o gpurun_out/${TAG}_backtest $PROF > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu done rc=$?"
# four consecutive launches of the fp16-pair GEMM (layer 2, layer 3, read-out, layer 1 of the next pass)
ncu --set full --clock-control none --import-source on -k regex:gemm_tc16_kernel -s 61 -c 4 -o gpurun_out/${TAG}_gemm_tc16 $PROF > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ncu gemm done rc=$?"
