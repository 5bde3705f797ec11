timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
timeout 100 python scripts/tl_sweep.py 2>&1 | tail -12
