mkdir -p gpurun_out
python scripts/kernel_timeline.py > gpurun_out/r3a_timeline.log 2>&1; echo rc=$?
sed -n 4,12p gpurun_out/r3a_timeline.log | cut -c1-150; grep forecast_span gpurun_out/r3a_timeline.log | cut -c1-200
python scripts/ab_bench.py one koopman_mpc_portfolio_rebalancing_b200/libkmpc.so 3 2>&1 | tail -1 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_forecast.py -m gpu -x -q > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3a_pytest.log
