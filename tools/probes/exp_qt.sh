timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_api.py tests/test_gpu_fused.py -x -q -k "qt or QT or c2 or multi or quant" 2>&1 | tail -3
for f in "" "--f32"; do timeout -s KILL 300 python bench.py --qt $f --noise 1.3 --slab-log2 28 --steps 20 --warmup 3 --no-cpu --no-e2e --no-outlier-leg --no-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('qt $f ms', d['ms_compress'], d['ms_decompress'], 'frac', d['roofline']['phases']['compress']['frac'], d['roofline']['phases']['decompress']['frac'])"; done
B2="python bench.py --qt --noise 1.3 --slab-log2 28 --steps 2 --warmup 1 --no-cpu --no-e2e --no-outlier-leg --no-configs"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/qt_launches.csv $B2 > /dev/null 2>&1
python - <<'P'
import csv,collections
rows=list(csv.reader(open('gpurun_out/qt_launches.csv'))); hdr=None; agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        n=r[hdr.index('Kernel Name')].split('(')[0]; v=float(r[hdr.index('Metric Value')])
        a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
for n,a in agg.items():
    if 'dctz' in n: print(f"   {n[:50]:50s} x{a[0]:3d} avg {a[1]/a[0]/1e3:9.1f} us")
P
