# round-1 (second half) measurement pass after the timelag / classify / expand rework
mkdir -p gpurun_out
timeout 600 python scripts/kernel_bench.py > gpurun_out/r1_kernel_bench.jsonl 2> gpurun_out/r1_kernel_bench.err; echo "kb exit $?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench_1gpu.json 2> gpurun_out/r1_bench_1gpu.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference.json 2>&1; echo "ref exit $?"
B="python bench.py --steps 2 --warmup 3 --atoms 3000 --no-cpu --e2e-atoms 64 --allpairs-vectors 30000 --allpairs-symmetric 8192"
timeout 600 $B > gpurun_out/r1_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches.csv $B > gpurun_out/r1_ncu_launches.log 2>&1
ls -la gpurun_out | grep r1_
# full captures of the all-pairs kernels (symmetric 12 288^2 call)
python scripts/prof_pairs.py f64 > gpurun_out/pp.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pairs_i8_kernel -s 1 -c 1 -f -o gpurun_out/r1_prof_pairs_i8_f64 python scripts/prof_pairs.py f64 > gpurun_out/pp_ncu64.log 2>&1
python scripts/prof_pairs.py f32 > gpurun_out/pp.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pairs_i8_kernel -s 1 -c 1 -f -o gpurun_out/r1_prof_pairs_i8_f32 python scripts/prof_pairs.py f32 > gpurun_out/pp_ncu32.log 2>&1
# classification kernel
python scripts/kernel_bench.py --only classify > gpurun_out/kb_cls.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:classify_tiled_kernel -s 2 -c 1 -f -o gpurun_out/r1_prof_classify python scripts/kernel_bench.py --only classify > gpurun_out/kb_cls_ncu.log 2>&1
ls -la gpurun_out | grep r1_prof
