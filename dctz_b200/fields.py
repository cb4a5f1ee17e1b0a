"""Deterministic synthetic fields of the shapes BASELINE.json names.

The reference ships no data (its tests download CESM/Hurricane/NYX files, README.md:71-77) and its
tools/rand-gen.c only emits 10 000 unseeded int32 values, so every benchmark / parity input is
generated here.  Fields are flattened row-major exactly as dctz-test.c:77-91 treats them
(only N = r1*r2*... matters to the codec).
"""
from __future__ import annotations

import numpy as np

SEED = 20261018


def cesm_like(ny: int = 1800, nx: int = 3600, dtype=np.float64, seed: int = SEED) -> np.ndarray:
    """C1/C2: smooth 2-D 'CESM-ATM-like' field in ~[0,1] (=> sf = 0.1) with small Gaussian noise."""
    rng = np.random.default_rng(seed)
    y = np.arange(ny, dtype=np.float64)[:, None]
    x = np.arange(nx, dtype=np.float64)[None, :]
    f = (0.5 + 0.35 * np.sin(2 * np.pi * 7 * x / nx) * np.cos(2 * np.pi * 5 * y / ny)
         + 0.1 * np.sin(x / 9.0 + y / 13.0))
    f = f + 0.002 * rng.standard_normal((ny, nx))
    return np.ascontiguousarray(f.astype(dtype).reshape(-1))


def hurricane_like(nz: int = 100, ny: int = 500, nx: int = 500, dtype=np.float32, seed: int = SEED) -> np.ndarray:
    """C3: 3-D 'Hurricane-like' field, range ~[-25, 25] (=> sf = 10)."""
    rng = np.random.default_rng(seed)
    z = np.arange(nz, dtype=np.float64)[:, None, None]
    y = np.arange(ny, dtype=np.float64)[None, :, None]
    x = np.arange(nx, dtype=np.float64)[None, None, :]
    f = 20.0 * np.sin(x / 40.0) * np.cos(y / 55.0) * np.exp(-z / 60.0) + 5.0 * np.sin((x + y + z) / 11.0)
    f = f + 0.05 * rng.standard_normal((nz, ny, nx))
    return np.ascontiguousarray(f.astype(dtype).reshape(-1))


def nyx_like(n: int = 512, dtype=np.float64, seed: int = SEED) -> np.ndarray:
    """C4: 3-D 'NYX-like' heavy-tailed positive field exp(1.5 g) (many outliers)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, n, n), dtype=dtype)
    y = np.arange(n, dtype=np.float64)[:, None]
    x = np.arange(n, dtype=np.float64)[None, :]
    base = np.sin(2 * np.pi * 3 * x / n) * np.cos(2 * np.pi * 2 * y / n)
    for k in range(n):  # slab by slab to bound host memory
        g = 0.6 * base + 0.4 * np.sin(2 * np.pi * (k / n) * 5 + x / 37.0) + 0.1 * rng.standard_normal((n, n))
        out[k] = np.exp(1.5 * g)
    return out.reshape(-1)


def _hash32(i: np.ndarray) -> np.ndarray:
    """32-bit integer mix (xorshift-multiply); exact in uint64 arithmetic, reproducible on device."""
    h = (i ^ (i >> np.uint64(16))) & np.uint64(0xFFFFFFFF)
    h = (h * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    h = h ^ (h >> np.uint64(15))
    h = (h * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    h = h ^ (h >> np.uint64(16))
    return h


def hash_field(start: int, count: int, dim: int = 2048, seed: int = SEED) -> np.ndarray:
    """C5: elements [start, start+count) of the dim^3 double field
    v(i) = 20 + 15 tri(x/dim) tri(y/dim) + 5 tri(z/dim) + 2^-10 (hash32(i ^ seed)/2^32 - 0.5),
    tri(t) = 1 - |2t - 1|.  Only exactly representable operations (dyadic rationals), so the device
    twin (dctz_gpu_fill_hash_field) reproduces it bit for bit and any chunk can be checked on the host."""
    i = np.arange(start, start + count, dtype=np.uint64)
    d = np.uint64(dim)
    x = (i % d).astype(np.float64) / dim
    y = ((i // d) % d).astype(np.float64) / dim
    z = (i // (d * d)).astype(np.float64) / dim
    tri = lambda t: 1.0 - np.abs(2.0 * t - 1.0)
    h = _hash32((i ^ np.uint64(seed)) & np.uint64(0xFFFFFFFF)).astype(np.float64)
    return 20.0 + 15.0 * (tri(x) * tri(y)) + 5.0 * tri(z) + (h / 4294967296.0 - 0.5) / 1024.0


def small_cases(dtype=np.float64, seed: int = 7):
    """Correctness-only inputs: ragged tails, degenerate statistics, bin edges."""
    rng = np.random.default_rng(seed)
    cases = {}
    cases["one_block"] = rng.standard_normal(64) * 3
    cases["three_blocks_tail37"] = np.cumsum(rng.standard_normal(229)) * 0.1 + 4.0
    cases["tail32"] = np.sin(np.arange(64800) / 50.0) * 7 + rng.standard_normal(64800) * 0.01
    cases["tail1"] = rng.standard_normal(64 * 5 + 1)
    cases["tail63_odd"] = rng.standard_normal(64 * 2 + 63) * 100
    cases["only_tail"] = rng.standard_normal(17) + 2
    cases["all_equal"] = np.full(640, 3.25)
    cases["sf_one"] = np.concatenate([rng.uniform(1.0, 9.9, 1000), [9.99]])  # max in (1,10] -> sf == 1.0
    cases["negative_only"] = -np.abs(rng.standard_normal(1000)) * 1e3 - 1.0
    cases["tiny_values"] = rng.standard_normal(2000) * 1e-7
    spike = np.zeros(64 * 20) + 0.5
    spike[700] = 9.0
    cases["single_spike"] = spike
    cases["heavy_outliers"] = rng.standard_normal(64 * 300) * 5.0  # white noise: most AC are outliers
    cases["smooth"] = np.cos(np.arange(64 * 200) / 400.0) * 2.0
    return {k: np.ascontiguousarray(v.astype(dtype)) for k, v in cases.items()}
