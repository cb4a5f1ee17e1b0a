#!/usr/bin/env python3
"""BASELINE config[3]: NYX-shaped 512^3 double field, DCT via FP64 DMMA vs the register butterfly.
Times dctz_gpu_dct64_dev (transform only: read 8 B + write 8 B per element) for both variants with CUDA
events and prints one JSON line; run it under ncu to get the FP64 / tensor pipe utilisation of each kernel."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import dctz_b200
    from dctz_b200 import DOUBLE

    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = 512 ** 3
    ctx = dctz_b200.Context(0)
    g = torch.Generator(device="cuda").manual_seed(7)
    t = torch.arange(n, device="cuda", dtype=torch.float64)
    x = torch.exp(1.5 * (0.6 * torch.sin(t / 97.0) * torch.cos(t / 1013.0) + 0.1 * torch.randn(n, generator=g, device="cuda", dtype=torch.float64)))
    del t
    out = torch.empty_like(x)
    st = torch.cuda.current_stream()
    peak = 6471.1
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    res = {}
    ref = None
    flop = {0: 9.25, 1: 128.0, 2: 64.0}  # FP64 operations per element: generated butterfly / full matrix / even-odd split
    for name, variant in (("butterfly", 0), ("dmma", 1), ("dmma_even_odd_split", 2)):
        for _ in range(3):
            ctx.dct64_dev(x.data_ptr(), out.data_ptr(), n // 64, DOUBLE, False, variant, st.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(st)
        for _ in range(steps):
            ctx.dct64_dev(x.data_ptr(), out.data_ptr(), n // 64, DOUBLE, False, variant, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        gbs = 16.0 * n / ms / 1e6
        res[name] = dict(ms=ms, hbm_gbs=gbs, frac_of_measured_peak=gbs / peak, gelem_s=n / ms / 1e6,
                         fp64_flop_per_element=flop[variant], tflops=flop[variant] * n / ms / 1e9)
        if ref is None:
            ref = out.clone()
        else:
            res[name]["max_abs_diff_vs_butterfly"] = float((out - ref).abs().max().item())
            res["max_abs_coefficient"] = float(ref.abs().max().item())
    res["fp64_rate_tflops"] = dict(vector_dfma=ctx.fp64_rate(0), tensor_dmma_m8n8k4=ctx.fp64_rate(1),
                                   note="measured issue-rate probes (dctz_gpu_fp64_rate): independent DFMA chains / back-to-back mma.sync.m8n8k4.f64")
    print(json.dumps(dict(workload="config[3]: 512^3 double (NYX-like), forward 64-point DCT only, 16 B/element", peak_gbs=peak, **res)))


if __name__ == "__main__":
    main()
