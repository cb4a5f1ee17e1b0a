"""diagnostic: where do GPU bins differ from the oracle on the hurricane field at small error bounds?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dctz_b200
from dctz_b200 import fields, binding, FLOAT, DOUBLE
from tests import reflib

ctx = dctz_b200.Context(0)
x = fields.hurricane_like()
n = x.size
d = torch.from_numpy(x).cuda()
s = torch.cuda.current_stream().cuda_stream
for eb in (1e-3, 1e-4, 1e-5):
    o = reflib.oracle_compress(x, eb, False, want_coef=False)
    runs = []
    for rep in range(3):
        bins = torch.empty(n, dtype=torch.uint8, device="cuda"); dc = torch.empty(n // 64, dtype=torch.float32, device="cuda")
        ac = torch.empty(n, dtype=torch.float32, device="cuda"); info = torch.zeros(binding.INFO_BYTES, dtype=torch.uint8, device="cuda")
        ctx.compress_field_dev(d.data_ptr(), n, FLOAT, eb, False, bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), 0, 0, info.data_ptr(), s)
        torch.cuda.synchronize()
        runs.append(bins.cpu().numpy())
    i = binding.GpuInfo.from_buffer_copy(info.cpu().numpy().tobytes()).as_dict()
    for r, b in enumerate(runs):
        diff = np.nonzero(b != o["bin_index"])[0]
        blocks = np.unique(diff // 64)
        heavy = [int(bk) for bk in blocks if np.count_nonzero(diff // 64 == bk) > 8][:4000]
        print(f"eb={eb:g} run{r}: mismatches={diff.size} blocks={blocks.size} heavy_blocks={len(heavy)} n_out gpu={i['n_outliers']} oracle={o['ac'].size} same_as_run0={np.array_equal(b, runs[0])}")
        if heavy and r == 0:
            hb = np.array(heavy)
            print("  first heavy blocks:", hb[:20], " tile:", (hb // 32)[:20], " lane:", (hb % 32)[:20])
            print("  lane histogram:", np.bincount(hb % 32, minlength=32))
            print("  tile mod 4 histogram:", np.bincount((hb // 32) % 4, minlength=4), " blocks per heavy tile:", np.bincount(np.unique(hb // 32, return_counts=True)[1])[:34])
            bk = heavy[0]
            print("  gpu ", b[bk * 64: bk * 64 + 24]); print("  orcl", o["bin_index"][bk * 64: bk * 64 + 24])
            # does the garbage block equal the bins of some other block?
            for other in (bk - 32, bk + 32, bk - 1, bk + 1):
                if 0 <= other < n // 64:
                    print("   == oracle block", other, np.array_equal(b[bk * 64: bk * 64 + 64], o["bin_index"][other * 64: other * 64 + 64]))
    # host API path too
    g = ctx.compress_core(x, eb)
    diff = np.nonzero(g["bin_index"] != o["bin_index"])[0]
    print(f"eb={eb:g} host api: mismatches={diff.size}")
