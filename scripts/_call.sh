mkdir -p gpurun_out
timeout 300 python scripts/cfg3_run.py build_ab/libkmpc_prof10cl.so 7=1 > gpurun_out/r2v_cfg3_cluster_prof.log 2>&1; echo "cluster rc=$?"; grep -v -i warn gpurun_out/r2v_cfg3_cluster_prof.log | grep "block 0" | tail -2 | cut -c1-700
