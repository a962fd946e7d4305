bash scripts/gpu_check.sh r2x
