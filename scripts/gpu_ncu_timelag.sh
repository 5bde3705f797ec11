mkdir -p gpurun_out
B="timeout 300 python bench.py --steps 3 --warmup 3 --atoms 3000 --no-cpu --e2e-atoms 64"
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:timelag_kernel -s 2 -c 1 -o gpurun_out/prof_timelag_r1b $B > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
