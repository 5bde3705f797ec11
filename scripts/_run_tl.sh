timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "classif or fuzz" 2>&1 | tail -4
timeout 200 python scripts/kernel_bench.py --only classify 2>&1 | tail -8
