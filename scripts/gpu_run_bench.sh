mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "exit $?" >> gpurun_out/bench_full.err
tail -3 gpurun_out/bench_full.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print('value %.4e pairs/s  ms/step %.2f  launches %d' % (d['value'], d['ms_per_step'], d['gpu_launches']))
print('roofline', {k: d['roofline'][k] for k in ('achieved', 'frac', 'kernel_ms', 'kernel_share_of_step')})
print('single', d['roofline']['single_window_calls'])
print('e2e', d['e2e'])
print('clocks', d['clocks'])
print('cpu', d['cpu_baseline'])
print('allpairs', json.dumps(d['allpairs'], indent=1)[:1800])
PY
