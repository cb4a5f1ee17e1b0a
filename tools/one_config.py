"""run one BASELINE config through the single-field entry points a few times (ncu target)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dctz_b200
from dctz_b200 import fields, binding, FLOAT, DOUBLE

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
eb = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
make, code, qt = {"c1": (lambda: fields.cesm_like(), DOUBLE, False), "c2": (lambda: fields.cesm_like(dtype=np.float32), FLOAT, True),
                  "c3": (lambda: fields.hurricane_like(), FLOAT, False), "c4": (lambda: fields.nyx_like(), DOUBLE, False)}[name]
ctx = dctz_b200.Context(0)
s = torch.cuda.current_stream().cuda_stream
x = torch.from_numpy(make()).cuda()
n = x.numel()
bins = torch.empty(n, dtype=torch.uint8, device="cuda"); dc = torch.empty(n // 64, dtype=torch.float32, device="cuda")
ac = torch.empty(n, dtype=torch.float32, device="cuda"); q = torch.zeros(64, dtype=x.dtype, device="cuda"); qr = torch.zeros(64, dtype=x.dtype, device="cuda")
info = torch.zeros(binding.INFO_BYTES, dtype=torch.uint8, device="cuda"); o = torch.empty_like(x)
for rep in range(reps):
    ctx.compress_field_dev(x.data_ptr(), n, code, eb, qt, bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), q.data_ptr(), qr.data_ptr(), info.data_ptr(), s)
    torch.cuda.synchronize()
    i = binding.GpuInfo.from_buffer_copy(info.cpu().numpy().tobytes()).as_dict()
    ctx.decompress_dev(bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), i["n_outliers"], q.data_ptr() if qt else 0, n, code, eb, i["sf"], qt, o.data_ptr(), s)
    torch.cuda.synchronize()
print(name, "ok", i["n_outliers"] / n)
