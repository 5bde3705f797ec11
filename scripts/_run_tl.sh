timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python scripts/kernel_bench.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['kernel'].ljust(22), d['config'][-44:].ljust(44), d.get('frac_hbm'), d.get('tflops_equiv', ''))
"
