# A/B of two builds of libdctz_gpu.so on the SAME box: dctz_b200/bin/libA.so, libB.so; alternating runs
for round in 1 2 3; do for v in A B; do cp dctz_b200/bin/lib$v.so dctz_b200/libdctz_gpu.so
  for a in "" 0; do DCTZ_DECOMP_AHEAD=$a timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); o=d['outlier_leg']; print('$v ahead=[$a] dec', round(d['ms_decompress'],4), 'comp', round(d['ms_compress'],4), 'outlier dec ec/qt/f32/f32qt', round(o['ms_decompress'],4), round(o['qt_mode']['ms_decompress'],4), round(o['f32']['ms_decompress'],4), round(o['f32_qt']['ms_decompress'],4))"; done; done; done
cp dctz_b200/bin/libA.so dctz_b200/libdctz_gpu.so
