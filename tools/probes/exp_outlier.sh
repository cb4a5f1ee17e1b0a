timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_api.py tests/test_gpu_fused.py -x -q 2>&1 | tail -3
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null > gpurun_out/s3i.json; python - <<'P'
import json
d=json.load(open('gpurun_out/s3i.json'))
print('value', d['value'], d['ms_compress'], d['ms_decompress'])
o=d['outlier_leg']
for k,v in (('ec',o),('qt',o['qt_mode']),('f32',o['f32']),('f32_qt',o['f32_qt'])):
    print(k, 'ms', round(v['ms_compress'],4), round(v['ms_decompress'],4), 'frac', round(v['compress_frac'],3), round(v['decompress_frac'],3))
for k,v in d['configs'].items():
    print(k, 'ms', round(v['ms_compress'],4), round(v['ms_decompress'],4), 'frac', round(v['compress_frac'],3), round(v['decompress_frac'],3), 'p', round(v['outlier_fraction'],3), v.get('parity_window',{}).get('ties'))
P
