timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for a in 0 1; do DCTZ_DECOMP_AHEAD=$a timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); o=d['outlier_leg']; print('ahead $a: value', d['value'], 'ms_decompress', d['ms_decompress'], 'ms_compress', d['ms_compress'], 'outlier leg dec ms', o['ms_decompress'], o['qt_mode']['ms_decompress'], o['f32']['ms_decompress'], o['f32_qt']['ms_decompress'], 'comp', o['ms_compress'])"; done
for a in 0 1; do DCTZ_DECOMP_AHEAD=$a timeout -s KILL 300 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c4 ahead $a: ms_decompress', d['ms_decompress'], 'ms_compress', d['ms_compress'])"; done
