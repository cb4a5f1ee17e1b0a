for h in 0 7 5 0 7; do
DCTZ_L2_HINTS=$h timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); o=d['outlier_leg']; print('hints $h: 2^30 dec', round(d['ms_decompress'],4), 'comp', round(d['ms_compress'],4), '| 2^28 5% dec ec/qt/f32/f32qt', round(o['ms_decompress'],4), round(o['qt_mode']['ms_decompress'],4), round(o['f32']['ms_decompress'],4), round(o['f32_qt']['ms_decompress'],4))"
DCTZ_L2_HINTS=$h timeout -s KILL 300 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('hints $h: c4 dec', round(d['ms_decompress'],4))"
done
DCTZ_L2_HINTS=7 timeout -s KILL 600 python -m pytest tests/test_gpu_device_api.py tests/test_gpu_parity.py -x -q 2>&1 | tail -2
