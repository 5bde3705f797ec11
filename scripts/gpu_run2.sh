mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
B="python bench.py --steps 3 --warmup 3 --atoms 3000 --no-cpu --e2e-atoms 64"
for cfg in "0 0" "6 1" "7 2" "8 2" "9 2" "12 1" "12 2" "12 3" "18 3"; do
  set -- $cfg
  echo "== split=$1 ahead=$2"
  SOAP_TL_SPLIT=$1 SOAP_TL_AHEAD=$2 $B 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.3e ms/step %.2f kernel_ms %.2f GB/s %.0f frac %.3f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['achieved'], r['frac']))
    elif 'rror' in l: print(l.strip())
"
done
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:timelag_kernel -s 2 -c 1 -o gpurun_out/prof_timelag_r1 $B > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
