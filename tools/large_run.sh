#!/bin/bash
# SURVEY.md §8f-2 on the box: a field beyond one stream's `int N` (2^31 + 323 floats = 8 GiB) through the drop-in CLI:
# DCTZMS01 container of block-aligned sub-streams that share the global sf, DCTZ_GPUS devices, then every sub-stream
# walked by dctz-dump and the reconstruction checked against the input.  Output: gpurun_out/large_run.log
set -u
G=${1:-2}
D=/tmp/dctz_large
mkdir -p $D gpurun_out
python - <<'P'
import numpy as np, time
n = (1 << 31) + 323
t0 = time.time()
rng = np.random.default_rng(7)
step = 1 << 26
t = np.arange(step, dtype=np.float32)
base = (20.0 + 15.0 * np.sin(t / 4099.0) * np.cos(t / 257.0) + 0.02 * rng.standard_normal(step).astype(np.float32)).astype(np.float32)
with open('/tmp/dctz_large/field.f32', 'wb') as f:  # the same 2^26-element pattern, shifted a little from repetition to repetition
    for k, a in enumerate(range(0, n, step)):
        (base[:min(step, n - a)] + np.float32(0.125 * (k % 7))).tofile(f)
print('generated', n, 'floats in %.1f s' % (time.time() - t0))
P
cd $D
( echo "== DCTZ_GPUS=$G dctz-ec-test -f 1E-3 var field.f32 $(( (1<<31) + 323 ))"; TIME=1 DCTZ_GPUS=$G DCTZ_NO_DUMPS=1 $OLDPWD/dctz_b200/bin/dctz-ec-test -f 1E-3 var field.f32 $(( (1<<31) + 323 )) 2>&1 | grep -v "^uncompressed\|^outSize" | tail -20; ls -la field.f32*; $OLDPWD/dctz_b200/bin/dctz-dump field.f32.ec.1E-3.zms 2>&1 | head -40 ) > $OLDPWD/gpurun_out/large_run.log 2>&1
python - <<'P' >> $OLDPWD/gpurun_out/large_run.log 2>&1
import numpy as np
n = (1 << 31) + 323
a = np.memmap('/tmp/dctz_large/field.f32', dtype=np.float32, mode='r', shape=(n,))
b = np.memmap('/tmp/dctz_large/field.f32.ec.1E-3.zms.r', dtype=np.float32, mode='r', shape=(n,))
m = 0.0
for s in range(0, n, 1 << 27):
    m = max(m, float(np.max(np.abs(a[s:s + (1 << 27)].astype(np.float64) - b[s:s + (1 << 27)]))))
print('max |reconstruction - input| over all %d elements: %.4g (sf 10, eb 1E-3: bound on the coefficients 1e-2)' % (n, m))
P
cd $OLDPWD; rm -rf $D
cat gpurun_out/large_run.log
