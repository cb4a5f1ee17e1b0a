"""PCIe link probe: H2D alone, D2H alone, both at once (pinned host memory, two streams), by chunk size.
Prints one JSON line.  Used to put the e2e leg's numbers against what the link can do."""
import json
import time

import torch

dev = torch.device("cuda", 0)
n = 1 << 30
h_up = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_dn = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_up = torch.empty(n, dtype=torch.uint8, device=dev)
d_dn = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, dn, chunk, reps=3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for off in range(0, n, chunk):
            if up:
                with torch.cuda.stream(s1):
                    d_up[off:off + chunk].copy_(h_up[off:off + chunk], non_blocking=True)
            if dn:
                with torch.cuda.stream(s2):
                    h_dn[off:off + chunk].copy_(d_dn[off:off + chunk], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * n / 1e9 / dt


out = {}
run(True, True, n, 1)
for chunk in (n, 64 << 20, 32 << 20, 16 << 20):
    out[f"chunk_{chunk >> 20}MiB"] = dict(h2d=run(True, False, chunk), d2h=run(False, True, chunk), both_each_direction=run(True, True, chunk))
print(json.dumps(out))
