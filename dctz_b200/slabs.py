"""Host-side slab logic for multi-GPU runs: one process per GPU, a field is cut into
contiguous slabs of whole 64-element blocks, and the ONLY data-path collective is the exchange of the
per-slab statistics {max|x|, min|x|, sum(x)} that the adaptive quantiser needs (the scaling factor is a
function of the global max, util.c:28).  QT mode adds one tiny max-reduction of the 64-entry qtable.
Outlier segments are concatenated in rank order by whoever assembles the stream; no collective."""
from __future__ import annotations

import math

BLK = 64


def partition(n_elements: int, world: int):
    """[(start, count)] per rank: contiguous runs of whole blocks, only the last rank may own the
    partial tail block.  Ranks beyond the data get (n, 0)."""
    nblk = (n_elements + BLK - 1) // BLK
    per = (nblk + world - 1) // world
    out = []
    for r in range(world):
        b0 = min(r * per, nblk)
        b1 = min((r + 1) * per, nblk)
        start = b0 * BLK
        end = min(b1 * BLK, n_elements)
        out.append((start, max(end - start, 0)))
    return out


def merge_stats(triples):
    """Reduce per-rank (max, min, sum) triples in RANK ORDER (deterministic sum), like k_finalize does
    on the device."""
    mx = max(t[0] for t in triples)
    mn = min(t[1] for t in triples)
    s = 0.0
    for t in triples:
        s += t[2]
    return mx, mn, s


_libm = None


def scaling_factor(max_abs: float, single: bool = False) -> float:
    """util.c:28 / util.c:42 evaluated with the host C libm (numpy's float32 log10 is a different
    implementation and disagrees at exact powers of ten); the device tables reproduce the same values."""
    global _libm
    if _libm is None:
        import ctypes

        _libm = ctypes.CDLL("libm.so.6")
        _libm.log10.restype = ctypes.c_double
        _libm.log10.argtypes = [ctypes.c_double]
        _libm.pow.restype = ctypes.c_double
        _libm.pow.argtypes = [ctypes.c_double, ctypes.c_double]
        _libm.log10f.restype = ctypes.c_float
        _libm.log10f.argtypes = [ctypes.c_float]
        _libm.powf.restype = ctypes.c_float
        _libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
    if single:
        return float(_libm.powf(10.0, math.ceil(_libm.log10f(max_abs)) - 1))
    return float(_libm.pow(10.0, math.ceil(_libm.log10(max_abs)) - 1))


def all_gather_stats(stats3, world: int):
    """torch.distributed all-gather of the 3 doubles of this rank (device tensor for NCCL, CPU tensor for
    gloo).  Returns a (3*world,) tensor laid out in rank order -- the `d_stats_all` argument of
    dctz_gpu_compress_dev."""
    import torch
    import torch.distributed as dist

    out = torch.empty(3 * world, dtype=stats3.dtype, device=stats3.device)
    if world == 1:
        out.copy_(stats3)
    else:
        dist.all_gather_into_tensor(out, stats3)
    return out


def last_rank_with_data(n_elements: int, world: int) -> int:
    """the rank whose slab holds the field's last block (trailing ranks of a small field may hold nothing)"""
    parts = partition(n_elements, world)
    return max(r for r, (_, c) in enumerate(parts) if c > 0)


def all_reduce_qtable(qraw, rank: int, world: int, src_last: int | None = None):
    """QT mode: entries 1..63 are per-position maxima of |outlier| -> MAX over ranks; entry 0 is the DC
    coefficient of the field's LAST block -> broadcast from the last rank that HOLDS DATA (`src_last`,
    see last_rank_with_data; default: the last rank)."""
    import torch.distributed as dist

    if world == 1:
        return qraw
    src = world - 1 if src_last is None else int(src_last)
    first = qraw[0:1].clone()
    dist.all_reduce(qraw, op=dist.ReduceOp.MAX)
    dist.broadcast(first, src=src)
    qraw[0:1] = first
    return qraw
