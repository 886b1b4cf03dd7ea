mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2> gpurun_out/bench_s.err | cut -c1-400
