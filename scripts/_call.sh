mkdir -p gpurun_out
for e in 1 2 1 2; do
python scripts/kernel_timeline.py embed=$e > gpurun_out/r3b_timeline_$e.log 2>&1; echo "embed=$e rc=$?"
sed -n 5,5p gpurun_out/r3b_timeline_$e.log | cut -c1-150; grep forecast_span gpurun_out/r3b_timeline_$e.log | cut -c1-120; sed -n 16,20p gpurun_out/r3b_timeline_$e.log | cut -c1-80
done
