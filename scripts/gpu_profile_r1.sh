# round-1 profiling pass: launch list of a short bench run + full captures of the two dominant kernels
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --atoms 3000 --no-cpu --e2e-atoms 64 --allpairs-vectors 30000"
timeout 600 $B > gpurun_out/r1_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv $B > gpurun_out/r1_ncu_launches.log 2>&1
timeout 600 $B > gpurun_out/r1_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:timelag_kernel -s 3 -c 1 -o gpurun_out/r1_prof_timelag $B > gpurun_out/r1_ncu_timelag.log 2>&1
timeout 600 $B > gpurun_out/r1_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pairs_i8_kernel -s 1 -c 1 -o gpurun_out/r1_prof_pairs_i8 $B > gpurun_out/r1_ncu_pairs.log 2>&1
tail -2 gpurun_out/r1_ncu_launches.log gpurun_out/r1_ncu_timelag.log gpurun_out/r1_ncu_pairs.log
ls -la gpurun_out | tail -12
