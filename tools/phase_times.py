"""where the microseconds of the small BASELINE configs go: phase stamps of the single-launch kernels"""
import os, sys, json
os.environ.setdefault("DCTZ_FUSED_STAMPS", "1")  # the stamps are off unless asked for
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dctz_b200
from dctz_b200 import fields, binding, FLOAT, DOUBLE

ctx = dctz_b200.Context(0)
s = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush2 = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")


def flush_l2():
    """write a buffer larger than L2, then read another one: L2 ends up full of CLEAN foreign lines (a dirty L2 would
    charge the next kernel for the write-back of the flush itself)"""
    flush.zero_()
    flush2.sum()


out = {}
for name, make, code, qt, eb in (("c1", lambda: fields.cesm_like(), DOUBLE, False, 1e-3), ("c2", lambda: fields.cesm_like(dtype=np.float32), FLOAT, True, 1e-3),
                                 ("c3", lambda: fields.hurricane_like(), FLOAT, False, 1e-3), ("c3@1E-5", lambda: fields.hurricane_like(), FLOAT, False, 1e-5)):
    x = torch.from_numpy(make()).cuda()
    n = x.numel()
    bins = torch.empty(n, dtype=torch.uint8, device="cuda"); dc = torch.empty(n // 64, dtype=torch.float32, device="cuda")
    ac = torch.empty(n, dtype=torch.float32, device="cuda"); q = torch.zeros(64, dtype=x.dtype, device="cuda"); qr = torch.zeros(64, dtype=x.dtype, device="cuda")
    info = torch.zeros(binding.INFO_BYTES, dtype=torch.uint8, device="cuda"); o = torch.empty_like(x)
    rows = []
    for rep in range(6):
        flush_l2()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        ctx.compress_field_dev(x.data_ptr(), n, code, eb, qt, bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), q.data_ptr(), qr.data_ptr(), info.data_ptr(), s)
        e[1].record()
        torch.cuda.synchronize()
        i = binding.GpuInfo.from_buffer_copy(info.cpu().numpy().tobytes()).as_dict()
        pc = ctx.fused_phase_times(0)
        flush_l2()
        e[1].record()
        ctx.decompress_dev(bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), i["n_outliers"], q.data_ptr() if qt else 0, n, code, eb, i["sf"], qt, o.data_ptr(), s)
        e[2].record()
        torch.cuda.synchronize()
        pd = ctx.fused_phase_times(1)
        rows.append((pc, pd))
    pc, pd = rows[-1]
    print(name, "compress stamps [start, sampled, past barrier 1, compressed, past barrier 2, end] (earliest, latest CTA) us:", [(round(a, 1), round(b, 1)) for a, b in pc])
    print(name, "decompress stamps us:", [(round(a, 1), round(b, 1)) for a, b in pd])
    out[name] = dict(compress=pc, decompress=pd)
json.dump(out, open("gpurun_out/phase_times.json", "w"))
