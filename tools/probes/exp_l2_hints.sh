B="python bench.py --slab-log2 26 --steps 20 --warmup 3 --no-cpu --no-e2e --no-outlier-leg --no-configs"
for h in 0 1 0 1; do DCTZ_L2_HINTS=$h $B 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('hints $h: ms_decompress', d['ms_decompress'], 'ms_compress', d['ms_compress'])"; done
for h in 0 1; do DCTZ_L2_HINTS=$h ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_decompress|k_count_bins" -s 8 -c 2 --csv python bench.py --slab-log2 26 --steps 2 --warmup 3 --no-cpu --no-e2e --no-outlier-leg --no-configs 2>/dev/null | grep -v "^==" | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r)>5 and r[0]!='ID' and 'k_' in r[4]: print('hints $h', r[4][:30], r[-3], r[-2], r[-1])"; done
