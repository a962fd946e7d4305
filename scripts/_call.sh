for d in randn relu zeros; do python scripts/gemm_shapes.py $d 32718,1024,1024 32718,250,1024 2>&1 | grep -v -i Warn; done
