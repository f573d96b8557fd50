mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_async_grid.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -8
bash tools/gpu_variants.sh r08 C2 0
bash tools/gpu_variants.sh r08 headline 0
