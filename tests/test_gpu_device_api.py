"""GPU: the device-resident C-ABI (what bench.py and a multi-GPU caller use) -- slabs with exchanged
statistics must reproduce the single-field result bit for bit; size-independent properties at a
BASELINE-sized field."""
import numpy as np
import pytest
import torch

from dctz_b200 import DOUBLE, FLOAT, binding, fields, slabs
from tests import parity, reflib

pytestmark = pytest.mark.gpu


def _info(t):
    return binding.GpuInfo.from_buffer_copy(t.cpu().numpy().tobytes()).as_dict()


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def compress_slabs(ctx, x, eb, qt, world):
    """Emulate `world` ranks on one GPU, one after the other (no inter-kernel waiting is involved)."""
    code = DOUBLE if x.dtype == np.float64 else FLOAT
    tdt = torch.float64 if code == DOUBLE else torch.float32
    parts = slabs.partition(x.size, world)
    s = torch.cuda.current_stream().cuda_stream
    xs = [_dev(x[a:a + c]) if c else None for a, c in parts]
    stats_all = torch.zeros(3 * world, dtype=torch.float64, device="cuda")
    for r, (a, c) in enumerate(parts):
        assert c > 0
        ctx.stats_dev(xs[r].data_ptr(), c, code, stats_all[3 * r:].data_ptr(), s)
    outs = []
    for r, (a, c) in enumerate(parts):
        nblk = (c + 63) // 64
        o = dict(bins=torch.empty(c, dtype=torch.uint8, device="cuda"), dc=torch.empty(nblk, dtype=torch.float32, device="cuda"),
                 ac=torch.empty(c, dtype=torch.float32, device="cuda"), qraw=torch.zeros(64, dtype=tdt, device="cuda"),
                 qt=torch.zeros(64, dtype=tdt, device="cuda"), info=torch.zeros(binding.INFO_BYTES, dtype=torch.uint8, device="cuda"))
        ctx.compress_dev(xs[r].data_ptr(), c, x.size, code, eb, qt, stats_all.data_ptr(), world, r == 0, o["bins"].data_ptr(),
                         o["dc"].data_ptr(), o["ac"].data_ptr(), o["qraw"].data_ptr(), o["info"].data_ptr(), s)
        outs.append(o)
    return parts, xs, stats_all, outs, code


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("world", [2, 3])
def test_slabs_reproduce_the_single_field_result_ec(ctx, dtype, world):
    x = fields.small_cases(dtype)["tail32"]
    x = np.concatenate([x, fields.small_cases(dtype)["heavy_outliers"] * 0.02 + 3]).astype(dtype)
    eb = 1e-3
    whole = ctx.compress_core(x, eb)
    parts, xs, stats_all, outs, code = compress_slabs(ctx, x, eb, False, world)
    torch.cuda.synchronize()
    infos = [_info(o["info"]) for o in outs]
    assert all(i["sf"] == whole["sf"] and i["status"] == 0 for i in infos)
    bins = np.concatenate([o["bins"].cpu().numpy() for o in outs])
    dc = np.concatenate([o["dc"].cpu().numpy() for o in outs])
    ac = np.concatenate([o["ac"][: i["n_outliers"]].cpu().numpy() for o, i in zip(outs, infos)])
    assert np.array_equal(bins, whole["bin_index"]) and np.array_equal(dc, whole["dc"]) and np.array_equal(ac, whole["ac"])
    assert infos[0]["max_abs"] == whole["info"]["max_abs"] and infos[-1]["min_abs"] == whole["info"]["min_abs"]
    assert abs(infos[0]["mean"] - whole["mean"]) <= 1e-6 * abs(whole["mean"]) + 1e-12
    # decompress slab by slab: each slab starts at its own outlier offset
    s = torch.cuda.current_stream().cuda_stream
    tdt = torch.float64 if code == DOUBLE else torch.float32
    rec = []
    for (a, c), o, i in zip(parts, outs, infos):
        out = torch.empty(c, dtype=tdt, device="cuda")
        ctx.decompress_dev(o["bins"].data_ptr(), o["dc"].data_ptr(), o["ac"].data_ptr(), i["n_outliers"], 0, c, code, eb, whole["sf"], False,
                           out.data_ptr(), s)
        rec.append(out.cpu().numpy())
    want = ctx.decompress_core(whole["bin_index"], whole["dc"], whole["ac"], x.size, dtype, eb, whole["sf"])
    assert np.array_equal(np.concatenate(rec), want)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("world", [2, 3])
def test_slabs_reproduce_the_single_field_result_qt(ctx, dtype, world):
    """QT mode over several slabs (dctz-comp-lib.c:355-372, 443-533): per-slab compress leaves the outliers un-rescaled,
    the per-position maxima are MAX-reduced over the slabs (entry 0 = DC of the field's last block, from the last slab),
    then every slab rescales with the global table.  One context per emulated rank (a context holds one call's
    scratch).  The concatenation must equal the single-field result bit for bit, table included."""
    import dctz_b200

    x = fields.small_cases(dtype)["tail32"]
    x = np.concatenate([fields.small_cases(dtype)["heavy_outliers"] * 0.02 + 3, x, fields.small_cases(dtype)["heavy_outliers"][:64 * 77 + 5] * 0.3 + 2]).astype(dtype)
    eb = 1e-3
    whole = ctx.compress_core(x, eb, qt=True)
    code = DOUBLE if dtype == np.float64 else FLOAT
    tdt = torch.float64 if code == DOUBLE else torch.float32
    s = torch.cuda.current_stream().cuda_stream
    parts = slabs.partition(x.size, world)
    ranks = [dctz_b200.Context(0) for _ in range(world)]
    try:
        xs = [_dev(x[a:a + c]) for a, c in parts]
        stats_all = torch.zeros(3 * world, dtype=torch.float64, device="cuda")
        for r, (a, c) in enumerate(parts):
            ranks[r].stats_dev(xs[r].data_ptr(), c, code, stats_all[3 * r:].data_ptr(), s)
        outs = []
        for r, (a, c) in enumerate(parts):
            o = dict(bins=torch.empty(c, dtype=torch.uint8, device="cuda"), dc=torch.empty((c + 63) // 64, dtype=torch.float32, device="cuda"),
                     ac=torch.empty(c, dtype=torch.float32, device="cuda"), qraw=torch.zeros(64, dtype=tdt, device="cuda"),
                     qt=torch.zeros(64, dtype=tdt, device="cuda"), info=torch.zeros(binding.INFO_BYTES, dtype=torch.uint8, device="cuda"))
            ranks[r].compress_dev(xs[r].data_ptr(), c, x.size, code, eb, True, stats_all.data_ptr(), world, r == 0, o["bins"].data_ptr(),
                                  o["dc"].data_ptr(), o["ac"].data_ptr(), o["qraw"].data_ptr(), o["info"].data_ptr(), s)
            outs.append(o)
        # the exchange all_reduce_qtable performs over NCCL: MAX of entries 1..63, entry 0 from the last rank
        g = torch.stack([o["qraw"] for o in outs]).max(dim=0).values
        g[0] = outs[-1]["qraw"][0]
        for r, o in enumerate(outs):
            o["qraw"].copy_(g)
            ranks[r].qt_finish_dev(code, eb, o["qraw"].data_ptr(), o["qt"].data_ptr(), o["ac"].data_ptr(), o["info"].data_ptr(), s)
        torch.cuda.synchronize()
        infos = [_info(o["info"]) for o in outs]
        assert all(i["sf"] == whole["sf"] and i["status"] == 0 and i["n_qt_dropped"] == 0 for i in infos)
        bins = np.concatenate([o["bins"].cpu().numpy() for o in outs])
        dc = np.concatenate([o["dc"].cpu().numpy() for o in outs])
        ac = np.concatenate([o["ac"][: i["n_outliers"]].cpu().numpy() for o, i in zip(outs, infos)])
        assert np.array_equal(bins, whole["bin_index"]) and np.array_equal(dc, whole["dc"])
        assert all(np.array_equal(o["qt"].cpu().numpy(), whole["qtable"]) for o in outs)
        assert np.array_equal(g.cpu().numpy(), whole["qtable_raw"])
        assert np.array_equal(ac, whole["ac"])
        rec = []
        for (a, c), o, i, rk in zip(parts, outs, infos, ranks):
            out = torch.empty(c, dtype=tdt, device="cuda")
            rk.decompress_dev(o["bins"].data_ptr(), o["dc"].data_ptr(), o["ac"].data_ptr(), i["n_outliers"], o["qt"].data_ptr(), c, code, eb,
                              whole["sf"], True, out.data_ptr(), s)
            rec.append(out.cpu().numpy())
        want = ctx.decompress_core(whole["bin_index"], whole["dc"], whole["ac"], x.size, dtype, eb, whole["sf"], qt=True, qtable=whole["qtable"])
        assert np.array_equal(np.concatenate(rec), want)
    finally:
        for rk in ranks:
            rk.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_single_slab_device_api_equals_host_api_qt(ctx, dtype):
    x = (fields.small_cases(dtype)["heavy_outliers"] * 0.05 + 2).astype(dtype)
    eb = 1e-3
    whole = ctx.compress_core(x, eb, qt=True)
    code = DOUBLE if dtype == np.float64 else FLOAT
    tdt = torch.float64 if code == DOUBLE else torch.float32
    s = torch.cuda.current_stream().cuda_stream
    d = _dev(x)
    n = x.size
    bins = torch.empty(n, dtype=torch.uint8, device="cuda")
    dc = torch.empty((n + 63) // 64, dtype=torch.float32, device="cuda")
    ac = torch.empty(n, dtype=torch.float32, device="cuda")
    q, qraw = torch.zeros(64, dtype=tdt, device="cuda"), torch.zeros(64, dtype=tdt, device="cuda")
    info = torch.zeros(binding.INFO_BYTES, dtype=torch.uint8, device="cuda")
    stats = torch.zeros(3, dtype=torch.float64, device="cuda")
    ctx.stats_dev(d.data_ptr(), n, code, stats.data_ptr(), s)
    ctx.compress_dev(d.data_ptr(), n, n, code, eb, True, stats.data_ptr(), 1, True, bins.data_ptr(), dc.data_ptr(), ac.data_ptr(),
                     qraw.data_ptr(), info.data_ptr(), s)
    qraw = slabs.all_reduce_qtable(qraw, 0, 1)  # where the NCCL max-reduction sits when there are several ranks
    ctx.qt_finish_dev(code, eb, qraw.data_ptr(), q.data_ptr(), ac.data_ptr(), info.data_ptr(), s)
    torch.cuda.synchronize()
    i = _info(info)
    assert np.array_equal(bins.cpu().numpy(), whole["bin_index"]) and np.array_equal(dc.cpu().numpy(), whole["dc"])
    assert np.array_equal(ac[: i["n_outliers"]].cpu().numpy(), whole["ac"]) and np.array_equal(q.cpu().numpy(), whole["qtable"])
    o = reflib.oracle_compress(x, eb, True)
    # the table holds max |outlier| per position: the oracle's value up to the DCT tolerance
    assert np.allclose(q.cpu().numpy()[1:], o["qtable"][1:], rtol=parity.RTOL[np.dtype(dtype)] * 8, atol=0)


def test_hash_field_device_twin_is_bit_exact(ctx):
    n = 1 << 20
    for start in (0, 12345 * 64, (1 << 32) + 64 * 7):
        d = torch.empty(n, dtype=torch.float64, device="cuda")
        ctx.fill_hash_field(d.data_ptr(), start, n, 2048, fields.SEED, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy(), fields.hash_field(start, n, 2048, fields.SEED))


def test_full_size_properties_nyx_like(ctx):
    """config[3]-sized field (512^3 doubles, 1 GiB) generated on the device: round trip honours the bound
    per coefficient, the outlier count is consistent, compress is deterministic (two runs, same bytes),
    and a 2^20-element window agrees with the oracle."""
    n = 1 << 27
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(7)
    t = torch.arange(n, device="cuda", dtype=torch.float64)
    x = torch.exp(1.5 * (0.6 * torch.sin(t / 97.0) * torch.cos(t / 1013.0) + 0.1 * torch.randn(n, generator=g, device="cuda", dtype=torch.float64)))
    del t
    eb = 1e-3
    bufs = []
    for _ in range(2):
        bins = torch.empty(n, dtype=torch.uint8, device="cuda")
        dc = torch.empty(n // 64, dtype=torch.float32, device="cuda")
        ac = torch.empty(n, dtype=torch.float32, device="cuda")
        info = torch.zeros(binding.INFO_BYTES, dtype=torch.uint8, device="cuda")
        ctx.compress_field_dev(x.data_ptr(), n, DOUBLE, eb, False, bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), 0, 0, info.data_ptr(), s)
        torch.cuda.synchronize()
        bufs.append((bins, dc, ac, _info(info)))
    (b0, d0, a0, i0), (b1, d1, a1, i1) = bufs
    k = i0["n_outliers"]
    assert i0 == i1 and torch.equal(b0, b1) and torch.equal(d0, d1) and torch.equal(a0[:k], a1[:k])
    pos = torch.arange(n, device="cuda") % 64
    assert int(((b0 == 255) & (pos != 0)).sum()) == k and bool(torch.all(b0[::64] == 255))
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    ctx.decompress_dev(b0.data_ptr(), d0.data_ptr(), a0.data_ptr(), k, 0, n, DOUBLE, eb, i0["sf"], False, out.data_ptr(), s)
    torch.cuda.synchronize()
    err = ((out - x).abs().max() / i0["sf"]).item()
    assert err <= eb * (1 + 63 * np.sqrt(2)) / 8 * 1.01
    # a window against the oracle (same sf because the window is compressed with the field's statistics? no:
    # the oracle computes its own; pick the window holding the global maximum so that sf agrees)
    w0 = (int(torch.argmax(x.abs()).item()) // (1 << 20)) * (1 << 20)
    xw = x[w0:w0 + (1 << 20)].cpu().numpy()
    o = reflib.oracle_compress(xw, eb, False)
    assert o["stat"]["sf"] == i0["sf"]
    gb = b0[w0:w0 + (1 << 20)].cpu().numpy()
    diff = np.nonzero(gb != o["bin_index"])[0]
    if diff.size:
        dist = parity.boundary_distance(o["coef"][diff], eb, np.float64)
        assert np.all(dist <= 1e-12 * np.maximum(parity.block_max(o["coef"])[diff], 1e-300))
    assert np.array_equal(d0[w0 // 64:(w0 + (1 << 20)) // 64].cpu().numpy(), o["dc"]) or np.allclose(
        d0[w0 // 64:(w0 + (1 << 20)) // 64].cpu().numpy(), o["dc"], rtol=2e-7)


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("inverse", [False, True])
def test_transform_only_kernels_butterfly_and_dmma(ctx, variant, inverse):
    """dctz_gpu_dct64_dev: the register butterfly (variant 0), the FP64-DMMA matrix form (variant 1) and the matrix form with the
    even/odd split (variant 2) compute the same orthonormal DCT-II / DCT-III as the oracle (dct.c:55-103 / 115-205) to 1e-12."""
    rng = np.random.default_rng(4)
    nblk = 32 * 50 + 7  # a partial last tile
    x = rng.standard_normal(nblk * 64) * rng.choice([1e-3, 1.0, 30.0], nblk * 64)
    d = _dev(x)
    out = torch.empty_like(d)
    ctx.dct64_dev(d.data_ptr(), out.data_ptr(), nblk, DOUBLE, inverse=inverse, variant=variant, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = out.cpu().numpy().reshape(-1, 64)
    pick = list(range(0, nblk, 37)) + [nblk - 1]
    want = np.stack([reflib.oracle_dct(x.reshape(-1, 64)[i], inverse=inverse) for i in pick])
    scale = np.max(np.abs(want), axis=1, keepdims=True)
    assert np.max(np.abs(got[pick] - want) / scale) <= 1e-12


def test_transform_only_float(ctx):
    rng = np.random.default_rng(5)
    nblk = 32 * 20 + 3
    x = rng.standard_normal(nblk * 64).astype(np.float32)
    d = _dev(x)
    ctx.dct64_dev(d.data_ptr(), d.data_ptr(), nblk, FLOAT, inverse=False, variant=0, stream=torch.cuda.current_stream().cuda_stream)  # in place
    torch.cuda.synchronize()
    got = d.cpu().numpy().reshape(-1, 64).astype(np.float64)
    want = np.stack([reflib.oracle_dct(b) for b in x.reshape(-1, 64)[:64]]).astype(np.float64)
    assert np.max(np.abs(got[:64] - want) / np.max(np.abs(want), axis=1, keepdims=True)) <= 1e-5


def test_more_elements_than_int32(ctx):
    """The C-ABI takes size_t counts: a slab of 2^31 + 64*5 + 3 doubles (16 GiB; beyond the reference's `int N`)
    goes through compress and decompress on the device; every index computation must be 64-bit clean.
    Checked by properties: block markers, outlier count consistency, the error bound, and an oracle
    comparison of the LAST 2^16 elements (the partial tail block included)."""
    n = (1 << 31) + 64 * 5 + 3
    s = torch.cuda.current_stream().cuda_stream
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    ctx.fill_hash_field(x.data_ptr(), 0, n, 2048, fields.SEED, s)
    x[n - 40000] = -39.0  # an isolated spike far inside the last tiles: outliers with a 64-bit offset
    eb = 1e-3
    bins = torch.empty(n, dtype=torch.uint8, device="cuda")
    dc = torch.empty((n + 63) // 64, dtype=torch.float32, device="cuda")
    ac = torch.empty(1 << 24, dtype=torch.float32, device="cuda")  # this field has few outliers (checked below)
    info = torch.zeros(binding.INFO_BYTES, dtype=torch.uint8, device="cuda")
    # statistics first, to size AC_exact honestly: the API contract is "room for N floats", here we verify the count
    ctx.compress_field_dev(x.data_ptr(), n, DOUBLE, eb, False, bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), 0, 0, info.data_ptr(), s)
    torch.cuda.synchronize()
    i = _info(info)
    assert i["status"] == 0 and i["sf"] == 10.0 and 0 < i["n_outliers"] < (1 << 24)
    assert bool(torch.all(bins[::64] == 255))
    start = ((n - (1 << 16)) // 64) * 64
    pos = (torch.arange(start, n, device="cuda") % 64)
    k_tail = int(((bins[start:] == 255) & (pos != 0)).sum())
    assert k_tail > 0
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    ctx.decompress_dev(bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), i["n_outliers"], 0, n, DOUBLE, eb, i["sf"], False, out.data_ptr(), s)
    torch.cuda.synchronize()
    err = float(((out - x).abs().max() / i["sf"]).item())
    assert err <= eb * (1 + 63 * np.sqrt(2)) / 8 * 1.01
    # oracle on the last 2^16 elements, compressed as their own field with the same sf (max of the window is the spike)
    xw = x[start:].cpu().numpy()
    o = reflib.oracle_compress(xw, eb, False)
    assert o["stat"]["sf"] == i["sf"]
    assert np.array_equal(bins[start:].cpu().numpy(), o["bin_index"])
    assert np.allclose(dc[start // 64:].cpu().numpy(), o["dc"], rtol=2e-7, atol=0)
    assert np.allclose(ac[i["n_outliers"] - k_tail: i["n_outliers"]].cpu().numpy(), o["ac"], rtol=2e-7, atol=0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_compress_with_known_statistics_is_corrected_on_the_device(ctx, dtype):
    """dctz_gpu_compress_known_stats_dev: one read of the input, scaling factor from the caller's belief (here: the
    previous 'time step'); the true statistics are gathered while compressing and a wrong belief costs a second
    compress pass on the device -- the result is always the two-pass one."""
    code = DOUBLE if dtype == np.float64 else FLOAT
    s = torch.cuda.current_stream().cuda_stream
    x_prev = (fields.small_cases(dtype)["tail32"]).astype(dtype)            # max ~ 7.03 -> sf = 1
    n = x_prev.size
    d_prev = _dev(x_prev)
    stats = torch.zeros(3, dtype=torch.float64, device="cuda")
    ctx.stats_dev(d_prev.data_ptr(), n, code, stats.data_ptr(), s)

    def run(x):
        d = _dev(x)
        o = dict(bins=torch.empty(n, dtype=torch.uint8, device="cuda"), dc=torch.empty((n + 63) // 64, dtype=torch.float32, device="cuda"),
                 ac=torch.empty(n, dtype=torch.float32, device="cuda"), info=torch.zeros(binding.INFO_BYTES, dtype=torch.uint8, device="cuda"))
        ctx.compress_known_stats_dev(d.data_ptr(), n, n, code, 1e-3, False, stats.data_ptr(), 1, True, o["bins"].data_ptr(),
                                     o["dc"].data_ptr(), o["ac"].data_ptr(), 0, o["info"].data_ptr(), s)
        torch.cuda.synchronize()
        o["i"] = _info(o["info"])
        return o

    def same_as_two_pass(o, x, redone):
        want = ctx.compress_core(x, 1e-3)
        i = o["i"]
        assert i["status"] == 0 and i["n_exact_path"] == redone and i["sf"] == want["sf"], (i, want["sf"])
        assert i["max_abs"] == float(np.max(np.abs(x))) and i["min_abs"] == float(np.min(np.abs(x)))  # the TRUE extremes, not the caller's
        assert abs(i["sum"] - want["info"]["sum"]) <= 1e-9 * float(np.sum(np.abs(x.astype(np.float64)))) * (1 if dtype == np.float64 else 1e4)
        k = i["n_outliers"]
        assert np.array_equal(o["bins"].cpu().numpy(), want["bin_index"]) and np.array_equal(o["dc"].cpu().numpy(), want["dc"])
        assert k == want["ac"].size and np.array_equal(o["ac"][:k].cpu().numpy(), want["ac"])

    # (a) the next time step stays in the same decade: one pass, bit-identical to the two-pass result
    x_next = (x_prev * dtype(1.01)).astype(dtype)
    same_as_two_pass(run(x_next), x_next, 0)
    # (b) the field grew into the next decade / (c) shrank below it: corrected by a second pass, same result
    for factor in (20.0, 0.05):
        xx = (x_prev * dtype(factor)).astype(dtype)
        same_as_two_pass(run(xx), xx, 1)
    # (d) the maximum sits in the partial tail block only
    spike = x_next.copy()
    spike[-3] = dtype(55.0)
    same_as_two_pass(run(spike), spike, 1)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("qt", [False, True])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_single_read_path_equals_two_pass(ctx, dtype, qt, world):
    """sample -> [exchange] -> compress_spec -> [exchange] -> compress_spec_finish over 1..3 slabs (one context per
    emulated rank) against stats -> compress of the whole field: same bytes, same statistics.  Fields: one whose
    maximum every sample finds, one whose maximum is a single spike between the sample points of slab 1 (the belief is
    a decade too low: every slab is compressed twice), one of zeros with one value (the belief degenerates)."""
    import dctz_b200

    code = DOUBLE if dtype == np.float64 else FLOAT
    tdt = torch.float64 if code == DOUBLE else torch.float32
    s = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(5)
    n = 64 * 32 * 90 + 64 * 5 + 19
    base = (3.0 + np.sin(np.arange(n) / 41.0) * 2.0 + 0.02 * rng.standard_normal(n))
    spike = base.copy()
    spike[n // 2 + 777] = 83.0          # not on a sampled vector (the samples are 16 bytes every 4 KB)
    lonely = np.zeros(n)
    lonely[12345] = -37.0
    ranks = [dctz_b200.Context(0) for _ in range(world)]
    try:
        for name, field, redone in (("plain", base, 0), ("spike", spike, 1), ("lonely", lonely, 1)):
            x = field.astype(dtype)
            # reference: the two-pass device path on the whole field
            ref = _two_pass(ctx, x, code, tdt, 1e-3, qt, s)
            parts = slabs.partition(n, world)
            xs = [_dev(x[a:a + c]) for a, c in parts]
            belief_all = torch.zeros(3 * world, dtype=torch.float64, device="cuda")
            true_all = torch.zeros(3 * world, dtype=torch.float64, device="cuda")
            outs = []
            for r, (a, c) in enumerate(parts):
                ranks[r].sample_dev(xs[r].data_ptr(), c, code, belief_all[3 * r:].data_ptr(), s)
            for r, (a, c) in enumerate(parts):
                o = dict(bins=torch.empty(c, dtype=torch.uint8, device="cuda"), dc=torch.empty((c + 63) // 64, dtype=torch.float32, device="cuda"),
                         ac=torch.empty(c, dtype=torch.float32, device="cuda"), qraw=torch.zeros(64, dtype=tdt, device="cuda"),
                         qt=torch.zeros(64, dtype=tdt, device="cuda"), info=torch.zeros(binding.INFO_BYTES, dtype=torch.uint8, device="cuda"))
                ranks[r].compress_spec_dev(xs[r].data_ptr(), c, n, code, 1e-3, qt, belief_all.data_ptr(), world, r == 0, o["bins"].data_ptr(), o["dc"].data_ptr(),
                                           o["ac"].data_ptr(), o["qraw"].data_ptr(), o["info"].data_ptr(), true_all[3 * r:].data_ptr(), s)
                outs.append(o)
            for r, (a, c) in enumerate(parts):
                o = outs[r]
                ranks[r].compress_spec_finish_dev(xs[r].data_ptr(), c, n, code, 1e-3, qt, true_all.data_ptr(), world, r == 0, o["bins"].data_ptr(),
                                                  o["dc"].data_ptr(), o["ac"].data_ptr(), o["qraw"].data_ptr(), o["info"].data_ptr(), s)
            if qt:
                g = torch.stack([o["qraw"] for o in outs]).max(dim=0).values
                g[0] = outs[-1]["qraw"][0]
                for r, o in enumerate(outs):
                    o["qraw"].copy_(g)
                    ranks[r].qt_finish_dev(code, 1e-3, o["qraw"].data_ptr(), o["qt"].data_ptr(), o["ac"].data_ptr(), o["info"].data_ptr(), s)
            torch.cuda.synchronize()
            infos = [_info(o["info"]) for o in outs]
            assert all(i["status"] == 0 and i["sf"] == ref["i"]["sf"] and i["n_exact_path"] == redone for i in infos), (name, infos, ref["i"])
            assert all(i["max_abs"] == ref["i"]["max_abs"] and i["min_abs"] == ref["i"]["min_abs"] for i in infos), (name, infos[0], ref["i"])
            assert abs(infos[0]["sum"] - ref["i"]["sum"]) <= 1e-9 * max(1.0, float(np.sum(np.abs(x.astype(np.float64))))) * (1 if dtype == np.float64 else 1e4)
            bins = np.concatenate([o["bins"].cpu().numpy() for o in outs])
            dc = np.concatenate([o["dc"].cpu().numpy() for o in outs])
            ac = np.concatenate([o["ac"][: i["n_outliers"]].cpu().numpy() for o, i in zip(outs, infos)])
            assert np.array_equal(bins, ref["bins"]) and np.array_equal(dc, ref["dc"]) and np.array_equal(ac, ref["ac"]), name
            if qt:
                assert all(np.array_equal(o["qt"].cpu().numpy(), ref["qt"]) for o in outs), name
    finally:
        for rk in ranks:
            rk.close()


def _two_pass(ctx, x, code, tdt, eb, qt, s):
    n = x.size
    d = _dev(x)
    st = torch.zeros(3, dtype=torch.float64, device="cuda")
    o = dict(bins=torch.empty(n, dtype=torch.uint8, device="cuda"), dc=torch.empty((n + 63) // 64, dtype=torch.float32, device="cuda"),
             ac=torch.empty(n, dtype=torch.float32, device="cuda"), qraw=torch.zeros(64, dtype=tdt, device="cuda"), qt=torch.zeros(64, dtype=tdt, device="cuda"),
             info=torch.zeros(binding.INFO_BYTES, dtype=torch.uint8, device="cuda"))
    ctx.stats_dev(d.data_ptr(), n, code, st.data_ptr(), s)
    ctx.compress_dev(d.data_ptr(), n, n, code, eb, qt, st.data_ptr(), 1, True, o["bins"].data_ptr(), o["dc"].data_ptr(), o["ac"].data_ptr(), o["qraw"].data_ptr(),
                     o["info"].data_ptr(), s)
    if qt:
        ctx.qt_finish_dev(code, eb, o["qraw"].data_ptr(), o["qt"].data_ptr(), o["ac"].data_ptr(), o["info"].data_ptr(), s)
    torch.cuda.synchronize()
    i = _info(o["info"])
    return dict(i=i, bins=o["bins"].cpu().numpy(), dc=o["dc"].cpu().numpy(), ac=o["ac"][: i["n_outliers"]].cpu().numpy(), qt=o["qt"].cpu().numpy())


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("qt", [False, True])
@pytest.mark.parametrize("last_total", [1, 2, 3])
def test_decompress_mixed_outlier_density_and_unaligned_outlier_array(ctx, dtype, qt, last_total):
    """Tiles of every staging class in one field -- no outliers, a few, more than half of the coefficients, all of
    them -- and the outlier array at every 4-byte phase of a 16-byte granule (the kernel fetches the aligned
    superset of a tile's run and must clip it at both ends of the array)."""
    rng = np.random.default_rng(99)
    ntile = 40
    amp = np.repeat(np.concatenate([[0.0, 0.0], np.geomspace(1e-3, 3.0, ntile - 6), [0.0, 5.0, 0.0, 5.0]]), 2048)
    t = np.arange(amp.size)
    x = (2.0 + np.sin(t / 300.0) * (t >= 4096) + amp * rng.standard_normal(amp.size)).astype(dtype)  # tiles 0, 1: constant
    # the field ends with constant tiles that carry 3, then 2, then `last_total` isolated outliers (single DCT basis
    # functions): the last runs of the outlier array are shorter than a 16-byte granule, at every alignment
    n64 = np.arange(64)
    quiet = np.full(4 * 2048, 2.0)
    for b, ks in ((40, (5, 9, 33)), (64 + 9, (3, 17)), (96 + 31, (1, 2, 60)[:last_total])):
        for k in ks:
            quiet[b * 64:(b + 1) * 64] += 0.3 * np.cos(np.pi * (2 * n64 + 1) * k / 128.0)
    x = np.concatenate([x, quiet.astype(dtype), x[:37]])  # + ragged tail
    eb = 1e-4
    orc = reflib.oracle_compress(x, eb, qt, want_coef=False)
    sf = orc["stat"]["sf"]
    per_tile = (orc["bin_index"][: ntile * 2048].reshape(ntile, 2048) == 255).sum(axis=1) - 32
    assert per_tile.min() == 0 and 0 < np.sum((per_tile > 0) & (per_tile < 1000)) and per_tile.max() > 1900, per_tile
    tail_tiles = (orc["bin_index"][ntile * 2048:(ntile + 4) * 2048].reshape(4, 2048) == 255).sum(axis=1) - 32
    assert list(tail_tiles) == [0, 3, 2, last_total], tail_tiles
    want = reflib.oracle_decompress(orc["bin_index"], orc["dc"], orc["ac"], orc["qtable"], x.size, eb, sf, qt, np.dtype(dtype))
    code = DOUBLE if dtype == np.float64 else FLOAT
    tdt = torch.float64 if code == DOUBLE else torch.float32
    s = torch.cuda.current_stream().cuda_stream
    bins, dc = _dev(orc["bin_index"]), _dev(orc["dc"])
    qtab = _dev(orc["qtable"].astype(dtype)) if qt else None
    tol = (1e-12 if dtype == np.float64 else 1e-5) * float(np.max(np.abs(want))) * 8
    for phase in range(4):
        buf = torch.full((orc["ac"].size + 8,), float("nan"), dtype=torch.float32, device="cuda")  # NaN guards on both sides
        ac = buf[phase:phase + orc["ac"].size]
        ac.copy_(torch.from_numpy(orc["ac"]))
        out = torch.empty(x.size, dtype=tdt, device="cuda")
        ctx.decompress_dev(bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), orc["ac"].size, qtab.data_ptr() if qt else 0, x.size, code, eb, sf, qt,
                           out.data_ptr(), s)
        got = out.cpu().numpy()
        assert np.all(np.isfinite(got)), f"phase {phase}: a guard value leaked into the reconstruction"
        diff = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))))
        assert diff <= tol, (phase, diff, tol)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("qt", [False, True])
def test_decompress_fuzz_synthetic_streams(ctx, dtype, qt):
    """Decoder-only fuzz: synthetic (bin_index, DC, AC_exact) streams with a random outlier density per warp tile
    (none / a handful / half / all), random sizes around the tile and group boundaries, the outlier array at a random
    4-byte phase -- against the oracle's decoder (dctz-decomp-lib.c:358-511)."""
    rng = np.random.default_rng(4242 + (1 if qt else 0) + (2 if dtype == np.float32 else 0))
    code = DOUBLE if dtype == np.float64 else FLOAT
    tdt = torch.float64 if code == DOUBLE else torch.float32
    s = torch.cuda.current_stream().cuda_stream
    eb, sf = 1e-3, 10.0
    for case in range(24):
        nblk = int(rng.choice([1, 31, 32, 33, 64 * 32 - 1, 64 * 32 + 5, 3000, 5000]))
        rem = int(rng.choice([0, 0, 1, 37, 63]))
        n = nblk * 64 + rem
        ntile = (nblk + 31) // 32
        dens = rng.choice([0.0, 0.0, 2e-4, 2e-3, 0.05, 0.5, 0.9, 1.0], size=ntile)
        if case % 6 == 0:
            dens[:] = 0.0
            dens[rng.integers(ntile)] = 2e-3  # a single tile with a few outliers
        pm = np.repeat(dens, 2048)[: nblk * 64]
        pm = np.concatenate([pm, np.full(rem, 0.3)])
        bins = rng.integers(0, 255, size=n, dtype=np.uint8)
        bins[rng.random(n) < pm] = 255
        bins[::64] = 255  # the DC markers
        n_out = int(np.count_nonzero(bins == 255)) - (nblk + (1 if rem else 0))
        mag = 0.255 + rng.random(n_out) * 3.0
        acv = (mag * rng.choice([-1.0, 1.0], size=n_out)).astype(np.float32)
        dcv = (rng.standard_normal(nblk + (1 if rem else 0)) * 5).astype(np.float32)
        qtab = (1.0 + rng.random(64) * 20).astype(dtype) if qt else None
        want = reflib.oracle_decompress(bins, dcv, acv, qtab, n, eb, sf, qt, np.dtype(dtype))
        phase = int(rng.integers(4))
        buf = torch.full((n_out + 8,), float("nan"), dtype=torch.float32, device="cuda")
        ac = buf[phase:phase + n_out]
        if n_out:
            ac.copy_(torch.from_numpy(acv))
        out = torch.empty(n, dtype=tdt, device="cuda")
        d_bins, d_dc, d_qt = _dev(bins), _dev(dcv), (_dev(qtab) if qt else None)  # (kept alive across the call)
        ctx.decompress_dev(d_bins.data_ptr(), d_dc.data_ptr(), ac.data_ptr() if n_out else 0, n_out, d_qt.data_ptr() if qt else 0,
                           n, code, eb, sf, qt, out.data_ptr(), s)
        got = out.cpu().numpy()
        assert np.all(np.isfinite(got)), f"case {case}: a guard value leaked (n={n}, outliers={n_out}, phase={phase})"
        tol = (1e-12 if dtype == np.float64 else 1e-5) * float(np.max(np.abs(want))) * 8
        diff = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))))
        assert diff <= tol, (case, n, n_out, phase, diff, tol)
