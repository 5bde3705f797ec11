mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --atoms 1500 --no-cpu > gpurun_out/bench_small.log 2>&1; echo "exit $?" >> gpurun_out/bench_small.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "exit $?" >> gpurun_out/bench_full.log
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -c 1500 gpurun_out/bench_small.log; tail -c 3000 gpurun_out/bench_full.log
