B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs --no-outlier-leg"
for h in 3 0; do DCTZ_L2_HINTS=$h timeout -s KILL 600 ncu --cache-control none --set full --clock-control none -k regex:k_decompress -s 2 -c 1 -o gpurun_out/ahead_h$h -f $B > gpurun_out/ahead_ncu_$h.log 2>&1; tail -1 gpurun_out/ahead_ncu_$h.log; done
