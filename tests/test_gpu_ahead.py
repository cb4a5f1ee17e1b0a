"""GPU: the count-ahead decompress kernel (DCTZ_DECOMP_AHEAD=1: no k_count_bins / k_scan_groups pre-pass; every warp counts
the markers of its tiles ahead of processing them and looks its offsets up in three levels of published sums) must decode
exactly what the pre-pass path decodes -- small and ragged fields, slabs large enough for ticket batches of 4 tiles, dense
and sparse outliers, truncated outlier arrays, and two contexts decoding at the same time on one device."""
import os
import threading

import numpy as np
import pytest

import dctz_b200

pytestmark = pytest.mark.gpu


def _ctx(**env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return dctz_b200.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


@pytest.fixture(scope="module")
def prepass_ctx():
    c = _ctx(DCTZ_FUSED_MAX_MB=0, DCTZ_DECOMP_AHEAD=0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ahead_ctx():
    c = _ctx(DCTZ_FUSED_MAX_MB=0, DCTZ_DECOMP_AHEAD=1)
    yield c
    c.close()


def _field(n, dtype, noise, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)
    x = 3.0 + 2.0 * np.sin(t / 97.0) + 0.5 * np.cos(t / 5.3)
    if noise:
        x += noise * rng.standard_normal(n)
    return x.astype(dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("qt", [False, True])
@pytest.mark.parametrize("n,noise", [(64, 0.0), (64 * 33 + 17, 0.01), (64 * 32 * 7, 0.3), (64 * (32 * 300 + 5) + 63, 0.02), (64 * 32 * 5000, 0.05)])
def test_count_ahead_equals_prepass(prepass_ctx, ahead_ctx, dtype, qt, n, noise):
    x = _field(n, dtype, noise, seed=n % 89)
    g = prepass_ctx.compress_core(x, 1e-3, qt=qt)
    args = (g["bin_index"], g["dc"], g["ac"], n, dtype, 1e-3, g["sf"])
    l0 = ahead_ctx.launch_count
    a = ahead_ctx.decompress_core(*args, qt=qt, qtable=g.get("qtable"))
    launches = ahead_ctx.launch_count - l0
    b = prepass_ctx.decompress_core(*args, qt=qt, qtable=g.get("qtable"))
    assert np.array_equal(a, b)
    assert launches == (1 if n >= 64 else 0) + (1 if n % 64 else 0)  # the kernel alone (+ the tail block's)


def test_count_ahead_with_ticket_batches(prepass_ctx, ahead_ctx):
    """more than 16 tiles per resident warp: tickets stand for 4 consecutive tiles, units of 4 tiles are published"""
    n = 64 * 32 * 24000 + 64 * 3 + 11
    x = _field(n, np.float64, 0.08, seed=4)
    g = prepass_ctx.compress_core(x, 1e-3)
    assert g["info"]["n_outliers"] > 100000
    args = (g["bin_index"], g["dc"], g["ac"], n, np.float64, 1e-3, g["sf"])
    assert np.array_equal(ahead_ctx.decompress_core(*args), prepass_ctx.decompress_core(*args))


def test_count_ahead_reports_a_truncated_outlier_array(prepass_ctx, ahead_ctx):
    n = 64 * 32 * 600
    x = _field(n, np.float64, 0.2, seed=9)
    g = prepass_ctx.compress_core(x, 1e-3)
    short = g["ac"][: g["ac"].size // 2]
    for c in (ahead_ctx, prepass_ctx):
        with pytest.raises(dctz_b200.DctzGpuError):
            c.decompress_core(g["bin_index"], g["dc"], short, n, np.float64, 1e-3, g["sf"])
    r = ahead_ctx.decompress_core(g["bin_index"], g["dc"], g["ac"], n, np.float64, 1e-3, g["sf"])  # the context is still usable
    assert float(np.max(np.abs(r - x))) < 0.05


def test_two_contexts_decode_concurrently(ahead_ctx):
    """Every unit comes from the ticket counter, so a waiting warp only ever waits for warps that are running: two such
    kernels sharing the device (neither fully resident) must both finish."""
    other = _ctx(DCTZ_FUSED_MAX_MB=0, DCTZ_DECOMP_AHEAD=1)
    try:
        n = 64 * 32 * 12000
        xs = [_field(n, np.float64, 0.05, seed=s) for s in (1, 2)]
        gs = [ahead_ctx.compress_core(x, 1e-3) for x in xs]
        out = [None, None]

        def work(i, c):
            for _ in range(6):
                out[i] = c.decompress_core(gs[i]["bin_index"], gs[i]["dc"], gs[i]["ac"], n, np.float64, 1e-3, gs[i]["sf"])

        th = [threading.Thread(target=work, args=(i, c)) for i, c in enumerate((ahead_ctx, other))]
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=120)
        assert not any(t.is_alive() for t in th), "count-ahead decompress kernels of two contexts are stuck"
        for x, r in zip(xs, out):
            assert float(np.max(np.abs(r - x))) < 0.05
    finally:
        other.close()
