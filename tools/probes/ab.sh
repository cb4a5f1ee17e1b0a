# (WL="c1 c2 c3 c4" adds whole-field workloads; default c4)
# A/B of builds of libdctz_gpu.so on the SAME box: dctz_b200/bin/lib<V>.so for V in $@ (default A B); alternating runs
VARS=${@:-A B}
for round in $(seq ${ROUNDS:-2}); do for v in $VARS; do cp dctz_b200/bin/lib$v.so dctz_b200/libdctz_gpu.so
  timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); o=d['outlier_leg']; L=[o,o['qt_mode'],o['f32'],o['f32_qt']]
print('$v  c5 comp/dec', round(d['ms_compress'],4), round(d['ms_decompress'],4), ' 5% comp ec/qt/f32/f32qt', [round(x['ms_compress'],4) for x in L], 'dec', [round(x['ms_decompress'],4) for x in L])"
  for w in ${WL:-c4}; do timeout -s KILL 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v  $w comp/dec', round(d['ms_compress'],4), round(d['ms_decompress'],4))"
  done
done; done
cp dctz_b200/bin/libA.so dctz_b200/libdctz_gpu.so
