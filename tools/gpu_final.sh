mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -v "^sampled\|^Pair set\|^Congruent\|^Q size\|^num \|^object sym" | tail -15 > gpurun_out/r12_pytest_gpu.log; tail -5 gpurun_out/r12_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r12_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r12_smoke.log
timeout 600 python bench.py > gpurun_out/r12_bench_default.json 2> gpurun_out/r12_bench_default.err; echo "bench exit $?"; cut -c1-600 gpurun_out/r12_bench_default.json
