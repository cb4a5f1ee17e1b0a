#!/bin/bash
# GPU iteration loop: parity tests, then the headline / outlier / config timings (device-resident legs only)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2> gpurun_out/p1.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); ph=d['roofline']['phases']; o=d['outlier_leg']
print('c5 value %.0f comp %.3f ms dec %.3f ms  k_compress %.3f  phases c %.3f s %.3f d %.3f' % (d['value'], d['ms_compress'], d['ms_decompress'], d['roofline']['frac'], ph['compress']['frac'], ph['stats']['frac'], ph['decompress']['frac']))
print('outlier leg p=%.3f value %.0f comp %.3f dec %.3f' % (o['outlier_fraction'], o['value'], o['compress_frac'], o['decompress_frac']))"
for w in c1 c2 c3 c4; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); ph=d['roofline']['phases']
print('$w value %.0f comp %.4f ms dec %.4f ms  phases c %.3f d %.3f' % (d['value'], d['ms_compress'], d['ms_decompress'], ph['compress']['frac'], ph['decompress']['frac']))"; done
python bench.py --f32 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); ph=d['roofline']['phases']; o=d['outlier_leg']
print('f32 value %.0f comp %.3f ms dec %.3f ms  phases c %.3f d %.3f' % (d['value'], d['ms_compress'], d['ms_decompress'], ph['compress']['frac'], ph['decompress']['frac']))
print('f32 outlier leg', o and (o['outlier_fraction'], o['compress_frac'], o['decompress_frac']))"
python tools/realistic_field_exp.py x 2>&1 | tail -1
