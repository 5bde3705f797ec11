# round-1 (second half) measurement pass after the timelag / classify / expand rework
mkdir -p gpurun_out
timeout 600 python scripts/kernel_bench.py > gpurun_out/r1_kernel_bench.jsonl 2> gpurun_out/r1_kernel_bench.err; echo "kb exit $?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench_1gpu.json 2> gpurun_out/r1_bench_1gpu.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference.json 2>&1; echo "ref exit $?"
B="python bench.py --steps 2 --warmup 3 --atoms 3000 --no-cpu --e2e-atoms 64 --allpairs-vectors 30000 --allpairs-symmetric 8192"
timeout 600 $B > gpurun_out/r1_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches.csv $B > gpurun_out/r1_ncu_launches.log 2>&1
ls -la gpurun_out | grep r1_
