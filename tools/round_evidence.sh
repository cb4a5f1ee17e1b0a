#!/bin/bash
# Round evidence on ONE GPU: bench lines (default workload, reference arm, other configs), then the ncu launch list and
# one --set full capture of the same command (only after it has exited 0 without ncu).  Output: gpurun_out/ev_*.
set -u
R=${1:-r02}
O=gpurun_out
python bench.py > $O/ev_${R}_n1.json 2> $O/ev_${R}_n1.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > $O/ev_${R}_ref.json 2> $O/ev_${R}_ref.err
: > $O/ev_${R}_configs.jsonl
for w in c1 c2 c3 c4; do python bench.py --workload $w --steps 20 --warmup 3 --no-e2e >> $O/ev_${R}_configs.jsonl 2>/dev/null; done
python bench.py --f32 --steps 20 --warmup 3 --no-cpu --no-e2e >> $O/ev_${R}_configs.jsonl 2>/dev/null
python bench.py --f32 --noise 1.3 --slab-log2 28 --steps 20 --warmup 3 --no-cpu --no-e2e >> $O/ev_${R}_configs.jsonl 2>/dev/null
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-outlier-leg --no-configs"
$B > /dev/null 2>&1 || exit 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ev_${R}_launches.csv $B > /dev/null 2>&1
# per compress: k_compress (VERIFY instantiation) + k_compress (gate launch); per decompress: k_count_bins + k_decompress.
# 1 compress before the warm-up (2 matches) + 1 warm-up step (4): skip 6, capture the first timed step
ncu --set full --clock-control none --import-source on -k regex:"k_compress|k_decompress|k_count_bins" -s 6 -c 4 -o $O/ev_${R}_prof -f $B > $O/ev_${R}_ncu.log 2>&1
tail -2 $O/ev_${R}_ncu.log
# the 5 % outlier slab (EC): launch list + one capture of its compress / gather / decompress kernels
B2="python bench.py --noise 1.3 --slab-log2 28 --steps 2 --warmup 1 --no-cpu --no-e2e --no-outlier-leg --no-configs"
$B2 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ev_${R}_launches_outliers.csv $B2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_compress|k_decompress|k_gather_ec" -s 7 -c 4 -o $O/ev_${R}_prof_outliers -f $B2 > $O/ev_${R}_ncu2.log 2>&1
python tools/phase_times.py > $O/ev_${R}_phase_times.txt 2>&1
python tools/realistic_field_exp.py x 2>&1 | tail -1 > $O/ev_${R}_realistic.json
cat $O/ev_${R}_n1.json | cut -c1-400
