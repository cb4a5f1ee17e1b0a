# A/B of builds on the headline slab only (dctz_b200/bin/lib<V>.so for V in $@), alternating, ROUNDS rounds (default 2)
for round in $(seq ${ROUNDS:-2}); do for v in "$@"; do cp dctz_b200/bin/lib$v.so dctz_b200/libdctz_gpu.so
  timeout -s KILL 300 python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e --no-configs --no-outlier-leg 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v  c5 comp/dec', round(d['ms_compress'],4), round(d['ms_decompress'],4))"
done; done
cp dctz_b200/bin/lib$1.so dctz_b200/libdctz_gpu.so
