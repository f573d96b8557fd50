mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_async_grid.py tests/test_gpu_s4pcs.py -m gpu -q 2>&1 | tail -8
bash tools/gpu_variants.sh r11 C2 0
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r11_launches_C2.csv python bench.py --workload C2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r11_ncu_launch.log 2>&1; echo ncu exit $?
