mkdir -p gpurun_out
python scripts/gemm_shapes.py 9102,250,1024 32718,250,1024 32718,1024,1024 2>&1 | grep -v -i Warn
python scripts/kernel_timeline.py > gpurun_out/r2w_timeline.log 2>&1; echo rc=$?
sed -n 3,5p gpurun_out/r2w_timeline.log | cut -c1-150; sed -n 13,18p gpurun_out/r2w_timeline.log | cut -c1-150
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_forecast.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2w_pytest.log
