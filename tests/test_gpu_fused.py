"""GPU: the single-launch kernels for small fields (fused.cuh: one cooperative launch per direction, phases separated by
grid barriers) against the streaming multi-launch path (kernels.cuh) on the same inputs -- bit for bit -- and against
the oracle.  DCTZ_FUSED_MAX_MB=0 at context creation selects the streaming path."""
import os

import numpy as np
import pytest
import torch

import dctz_b200
from dctz_b200 import DOUBLE, FLOAT, binding, fields
from tests import parity, reflib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def streaming_ctx():
    old = os.environ.get("DCTZ_FUSED_MAX_MB")
    os.environ["DCTZ_FUSED_MAX_MB"] = "0"
    try:
        c = dctz_b200.Context(0)
    finally:
        if old is None:
            del os.environ["DCTZ_FUSED_MAX_MB"]
        else:
            os.environ["DCTZ_FUSED_MAX_MB"] = old
    yield c
    c.close()


def _field(n, dtype, kind, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    base = 3.0 + 2.5 * np.sin(t / 37.0) + 0.4 * np.cos(t / 3.3)
    if kind == "smooth":
        x = base + 1e-4 * rng.standard_normal(n)
    elif kind == "sparse":
        x = base + 0.01 * rng.standard_normal(n)
    else:  # dense: most coefficients are outliers
        x = base + 0.5 * rng.standard_normal(n)
    return x.astype(dtype)


def _launches(ctx, fn):
    l0 = ctx.launch_count
    r = fn()
    return r, ctx.launch_count - l0


SIZES = [64, 64 * 33, 64 * 32 * 7, 64 * (32 * 300 + 5), 64 * 32 * 1500]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("qt", [False, True])
@pytest.mark.parametrize("kind", ["smooth", "sparse", "dense"])
def test_single_launch_equals_streaming_path(ctx, streaming_ctx, dtype, qt, kind):
    for n in SIZES:
        x = _field(n, dtype, kind, seed=n % 97)
        eb = 1e-3 if kind != "dense" else 1e-4
        a, la = _launches(ctx, lambda: ctx.compress_core(x, eb, qt=qt))
        b, lb = _launches(streaming_ctx, lambda: streaming_ctx.compress_core(x, eb, qt=qt))
        assert la == 1 and lb >= 3, (la, lb)  # ONE launch for the whole compress direction
        for key in ("bin_index", "dc", "ac"):
            assert np.array_equal(a[key], b[key]), (n, key)
        if qt:
            assert np.array_equal(a["qtable"], b["qtable"]) and np.array_equal(a["qtable_raw"], b["qtable_raw"]), n
        ia, ib = a["info"], b["info"]
        for key in ("sf", "max_abs", "min_abs", "n_outliers", "status", "n_qt_dropped"):
            assert ia[key] == ib[key], (n, key, ia[key], ib[key])
        # (float data: the streaming path's single-read kernel adds a block's 64 values in float, like the reference's float sum)
        assert abs(ia["sum"] - ib["sum"]) <= (1e-9 if dtype == np.float64 else 1e-5) * np.sum(np.abs(x.astype(np.float64))) + 1e-300
        # decompress: both paths on the same stream
        ra, lda = _launches(ctx, lambda: ctx.decompress_core(a["bin_index"], a["dc"], a["ac"], n, dtype, eb, a["sf"], qt=qt, qtable=a.get("qtable")))
        rb, ldb = _launches(streaming_ctx, lambda: streaming_ctx.decompress_core(a["bin_index"], a["dc"], a["ac"], n, dtype, eb, a["sf"], qt=qt,
                                                                                 qtable=a.get("qtable")))
        assert lda == 1 and ldb >= 1, (lda, ldb)  # (streaming: count-ahead kernel alone, or pre-pass + scan + kernel)
        assert np.array_equal(ra, rb), n
    # and against the oracle at the largest size
    parity.check_compress(ctx, x, eb, qt)
    parity.check_decompress(ctx, x, eb, qt)


def test_barrier_bookkeeping_survives_failures_and_changing_grids(ctx):
    """The grid barrier counts arrivals in a counter that only grows; the host hands every launch its base value.  A
    degenerate field (all zeros: sf would be 0, util.c:28) leaves after the first barrier -- the arrivals it owes the
    second one must still be made, or every later launch would wait for ever.  Sizes alternate so that the grid changes."""
    rng = np.random.default_rng(1)
    for it in range(12):
        n = 64 * int(rng.choice([1, 40, 32 * 9, 32 * 400 + 3]))
        if it % 4 == 1:
            with pytest.raises(dctz_b200.DctzGpuError):
                ctx.compress_core(np.zeros(n), 1e-3)
            continue
        x = (rng.standard_normal(n) * 0.02 + 2).astype(np.float64 if it % 2 else np.float32)
        g = ctx.compress_core(x, 1e-3, qt=bool(it % 3 == 0))
        r = ctx.decompress_core(g["bin_index"], g["dc"], g["ac"], n, x.dtype, 1e-3, g["sf"], qt=bool(it % 3 == 0), qtable=g.get("qtable"))
        assert float(np.max(np.abs(r.astype(np.float64) - x))) < 0.02


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_single_launch_decoder_reports_corrupt_streams(ctx, dtype):
    x = _field(64 * 32 * 40, dtype, "sparse")
    g = ctx.compress_core(x, 1e-3)
    assert g["ac"].size > 1000
    for keep in (g["ac"].size - 1, g["ac"].size // 2, 0):
        with pytest.raises(dctz_b200.DctzGpuError) as e:
            ctx.decompress_core(g["bin_index"], g["dc"], g["ac"][:keep], x.size, dtype, 1e-3, g["sf"])
        assert e.value.code == -7
    ok = ctx.decompress_core(g["bin_index"], g["dc"], g["ac"], x.size, dtype, 1e-3, g["sf"])
    assert float(np.max(np.abs(ok - x))) < 0.05


def test_config_fields_take_the_single_launch_path(ctx):
    """BASELINE configs[0..2] (52 / 26 / 100 MB): one launch per direction, device-resident."""
    s = torch.cuda.current_stream().cuda_stream
    for make, code, qt in ((lambda: fields.cesm_like(), DOUBLE, False), (lambda: fields.cesm_like(dtype=np.float32), FLOAT, True),
                           (lambda: fields.hurricane_like(), FLOAT, False)):
        x = torch.from_numpy(make()).cuda()
        n = x.numel()
        bins = torch.empty(n, dtype=torch.uint8, device="cuda")
        dc = torch.empty(n // 64, dtype=torch.float32, device="cuda")
        ac = torch.empty(n, dtype=torch.float32, device="cuda")
        q, qraw = torch.zeros(64, dtype=x.dtype, device="cuda"), torch.zeros(64, dtype=x.dtype, device="cuda")
        info = torch.zeros(binding.INFO_BYTES, dtype=torch.uint8, device="cuda")
        out = torch.empty_like(x)
        l0 = ctx.launch_count
        ctx.compress_field_dev(x.data_ptr(), n, code, 1e-3, qt, bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), q.data_ptr(), qraw.data_ptr(), info.data_ptr(), s)
        torch.cuda.synchronize()
        i = binding.GpuInfo.from_buffer_copy(info.cpu().numpy().tobytes()).as_dict()
        l1 = ctx.launch_count
        ctx.decompress_dev(bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), i["n_outliers"], q.data_ptr() if qt else 0, n, code, 1e-3, i["sf"], qt, out.data_ptr(), s)
        torch.cuda.synchronize()
        assert l1 - l0 == 1 and ctx.launch_count - l1 == 1
        assert i["status"] == 0 and float((out - x).abs().max() / i["sf"]) <= 1e-3 * (1 + 63 * np.sqrt(2)) / 8 * 1.01 + (1e-3 if qt else 0)
