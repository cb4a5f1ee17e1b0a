import sys, json, numpy as np, torch
sys.path.insert(0, '/root/repo')
import dctz_b200
from dctz_b200 import DOUBLE, binding
ctx = dctz_b200.Context(0)
dev = torch.device('cuda')
ny = nx = 16384
# CESM-like smooth field + small noise, generated on device (same formula as fields.cesm_like scaled up)
y = torch.arange(ny, device=dev, dtype=torch.float64)[:, None]
x = torch.arange(nx, device=dev, dtype=torch.float64)[None, :]
f = 0.5 + 0.35 * torch.sin(2 * np.pi * 7 * x / 3600) * torch.cos(2 * np.pi * 5 * y / 1800) + 0.1 * torch.sin(x / 9.0 + y / 13.0)
g = torch.Generator(device=dev).manual_seed(1)
for noise in ((0.002,) if len(sys.argv) > 1 else (0.002, 0.0005, 0.005)):
    d = (f + noise * torch.randn(ny, nx, generator=g, device=dev, dtype=torch.float64)).reshape(-1).contiguous()
    n = d.numel()
    bins = torch.empty(n, dtype=torch.uint8, device=dev); dc = torch.empty(n // 64, dtype=torch.float32, device=dev)
    ac = torch.empty(n, dtype=torch.float32, device=dev); out = torch.empty_like(d)
    info = torch.zeros(binding.INFO_BYTES, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    def comp(): ctx.compress_field_dev(d.data_ptr(), n, DOUBLE, 1e-3, False, bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), 0, 0, info.data_ptr(), s)
    comp(); torch.cuda.synchronize()
    i = binding.GpuInfo.from_buffer_copy(info.cpu().numpy().tobytes()).as_dict()
    def dec(): ctx.decompress_dev(bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), 0, n, DOUBLE, 1e-3, i['sf'], False, out.data_ptr(), s)
    res = {}
    for name, fn in (('compress', comp), ('decompress', dec)):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 10
    p = i['n_outliers'] / n
    bc = (17.0625 + 4 * p) * n / res['compress'] / 1e6 / 6471.1
    bd = (9.0625 + 4 * p) * n / res['decompress'] / 1e6 / 6471.1
    print(json.dumps(dict(noise=noise, p=p, ms=res, compress_frac=bc, decompress_frac=bd, sf=i['sf'])))
