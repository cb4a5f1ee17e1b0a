"""GPU: the drop-in host library (libdctz_ec.so / libdctz_qt.so = DCTZ's public API over the GPU path):
stream layout, cross-decoding with the unmodified reference, legacy symbols, the CLI driver."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

from dctz_b200 import fields
from tests import parity, reflib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hostlib(qt):
    lib = C.CDLL(os.path.join(ROOT, "dctz_b200", f"libdctz_{'qt' if qt else 'ec'}.so"))
    lib.calc_psnr.restype = C.c_double
    return lib


def our_compress(x, eb, qt):
    lib = hostlib(qt)
    code = 1 if x.dtype == np.float64 else 0
    buf = np.array(x, copy=True)
    zbuf = np.zeros(2 * x.nbytes + 4096, np.uint8)
    var = reflib.TVar(code, eb, b"v", buf.ctypes.data)
    var_z = reflib.TVar(code, eb, b"v", zbuf.ctypes.data)
    out = C.c_size_t(0)
    with reflib._in_tmpdir():
        assert lib.dctz_compress(C.byref(var), C.c_int(x.size), C.byref(out), C.byref(var_z), C.c_double(eb)) == 1
        dumps = {n: np.fromfile(n, dtype=np.uint8) for n in ("bin_index.bin", "AC_exact.bin")}
    return zbuf[: out.value].copy(), buf, dumps


def our_decompress(stream, n, dtype, qt):
    lib = hostlib(qt)
    code = 1 if np.dtype(dtype) == np.float64 else 0
    z = np.array(stream, copy=True)
    r = np.zeros(n, dtype)
    var_z = reflib.TVar(code, 0.0, b"v", z.ctypes.data)
    var_r = reflib.TVar(code, 0.0, b"v", r.ctypes.data)
    assert lib.dctz_decompress(C.byref(var_z), C.byref(var_r)) == 1
    return r


def ref_decompress(stream, n, dtype, qt):
    lib = reflib.ref_lib(qt)
    code = 1 if np.dtype(dtype) == np.float64 else 0
    z = np.array(stream, copy=True)
    r = np.zeros(n, dtype)
    var_z = reflib.TVar(code, 0.0, b"v", z.ctypes.data)
    var_r = reflib.TVar(code, 0.0, b"v", r.ctypes.data)
    lib.dctz_decompress(C.byref(var_z), C.byref(var_r))
    return r


def split_stream(stream, dtype, qt):
    h = np.frombuffer(stream[:56].tobytes(), dtype=reflib.HEADER_DTYPE)[0]
    off, secs = 56, []
    for key in ("bindex_sz_compressed", "DC_sz_compressed", "AC_exact_sz_compressed"):
        secs.append(zlib.decompress(stream[off:off + int(h[key])].tobytes()))
        off += int(h[key])
    q = np.frombuffer(stream[off:].tobytes(), dtype=dtype) if qt else None
    assert stream.size == off + (64 * np.dtype(dtype).itemsize if qt else 0)
    return h, np.frombuffer(secs[0], np.uint8), np.frombuffer(secs[1], np.float32), np.frombuffer(secs[2], np.float32), q


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("qt", [False, True])
def test_stream_layout_and_contents(dtype, qt):
    x = fields.small_cases(dtype)["tail32"]
    eb = 1e-3
    stream, scaled, dumps = our_compress(x, eb, qt)
    h, bins, dc, ac, q = split_stream(stream, dtype, qt)
    o = reflib.oracle_compress(x, eb, qt)
    assert int(h["datatype"]) == (1 if dtype == np.float64 else 0) and int(h["num_elements"]) == x.size
    assert float(h["error_bound"]) == eb and int(h["tot_AC_exact_count"]) == ac.size
    sf = np.frombuffer(h["scaling_factor"].tobytes(), dtype=dtype)[0]
    assert float(sf) == o["stat"]["sf"]
    if qt:
        assert int(h["bindex_count"]) == x.size
    gpu = dict(bin_index=bins, dc=dc, ac=ac, qtable=q,
               info=dict(max_abs=o["stat"]["max"], min_abs=o["stat"]["min"], sf=float(sf), sum=o["stat"]["sum"], n_outliers=ac.size, n_qt_dropped=0))
    parity.compare_compress(gpu, o, x, eb, qt)
    assert np.array_equal(scaled, o["scaled"])  # the caller's buffer is left divided by sf, bit-exactly
    assert np.array_equal(dumps["bin_index.bin"], bins) and dumps["AC_exact.bin"].size == 4 * ac.size  # side files


@pytest.mark.skipif(not reflib.have_ref(), reason="oracle/_ref (the compiled reference) is not present")
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("qt", [False, True])
def test_cross_decoding_with_the_reference(dtype, qt):
    x = fields.small_cases(dtype)["three_blocks_tail37"]
    x = np.concatenate([x, fields.small_cases(dtype)["smooth"]])
    eb = 1e-3
    tol = parity.RTOL[np.dtype(dtype)] * 8 * float(np.max(np.abs(x)))
    ours, _, _ = our_compress(x, eb, qt)
    ref = reflib.ref_roundtrip(x, eb, qt)
    # our stream through the reference's decoder == our own decoder (to DCT tolerance)
    a = ref_decompress(ours, x.size, dtype, qt).astype(np.float64)
    b = our_decompress(ours, x.size, dtype, qt).astype(np.float64)
    assert np.max(np.abs(a - b)) <= tol
    # the reference's stream through our decoder == the reference's own reconstruction
    c = our_decompress(ref["stream"], x.size, dtype, qt).astype(np.float64)
    assert np.max(np.abs(c - ref["recon"].astype(np.float64))) <= tol
    # and the compressed size is the reference's, byte for byte, unless a tie moved a symbol
    assert abs(int(ours.size) - int(ref["stream"].size)) <= 16


def test_legacy_symbols():
    lib = hostlib(False)
    x = fields.small_cases(np.float64)["smooth"]
    # calc_data_stat (util.c:12-44)
    bs = (C.c_double * 5)()
    buf = np.array(x, copy=True)
    var = reflib.TVar(1, 0.0, b"v", buf.ctypes.data)
    lib.calc_data_stat(C.byref(var), bs, C.c_int(x.size))
    o = reflib.oracle_stat(x)
    assert bs[2] == o["max"] and bs[1] == o["min"] and bs[4] == o["sf"] and abs(bs[0] - o["mean"]) < 1e-9
    # gen_bins (binning.c:12-30)
    centre = np.zeros(255)
    lib.gen_bins(C.c_double(0), C.c_double(0), centre.ctypes.data_as(C.c_void_p), 255, C.c_double(1e-3))
    assert np.array_equal(centre, reflib.oracle_gen_bins(1e-3, np.float64))
    cf = np.zeros(255, np.float32)
    lib.gen_bins_f(C.c_float(0), C.c_float(0), cf.ctypes.data_as(C.c_void_p), 255, C.c_float(1e-3))
    assert np.array_equal(cf, reflib.oracle_gen_bins(1e-3, np.float32))
    # dct_init / dct_fftw / ifft_idct (dct.h:17-27)
    blk = np.ascontiguousarray(x[:64])
    co = np.zeros(64)
    back = np.zeros(64)
    lib.dct_init(64)
    lib.dct_fftw(blk.ctypes.data_as(C.c_void_p), co.ctypes.data_as(C.c_void_p), 64, 1)
    lib.dct_finish()
    lib.ifft_idct(64, co.ctypes.data_as(C.c_void_p), back.ctypes.data_as(C.c_void_p))
    lib.idct_finish()
    want = reflib.oracle_dct(blk)
    assert np.max(np.abs(co - want)) <= 1e-12 * np.max(np.abs(want)) and np.max(np.abs(back - blk)) <= 1e-12 * np.max(np.abs(blk))


@pytest.mark.parametrize("flavour,flag,dtype", [("ec", "-d", np.float64), ("qt", "-f", np.float32)])
def test_cli_round_trip(tmp_path, flavour, flag, dtype):
    x = fields.cesm_like(180, 360, dtype=dtype)
    src = tmp_path / "field.bin"
    x.tofile(src)
    exe = os.path.join(ROOT, "dctz_b200", "bin", f"dctz-{flavour}-test")
    p = subprocess.run([exe, flag, "1E-3", "var", str(src), "360", "180"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "total number of elements = 64800" in p.stdout and "CR = " in p.stdout and "done" in p.stdout
    z = np.fromfile(str(src) + f".{flavour}.1E-3.z", dtype=np.uint8)
    r = np.fromfile(str(src) + f".{flavour}.1E-3.z.r", dtype=dtype)
    assert r.size == x.size and z.size < x.nbytes
    o = reflib.oracle_compress(x, 1e-3, flavour == "qt")
    want = reflib.oracle_decompress(o["bin_index"], o["dc"], o["ac"], o["qtable"], x.size, 1e-3, o["stat"]["sf"], flavour == "qt", dtype)
    assert np.max(np.abs(r.astype(np.float64) - want.astype(np.float64))) <= parity.RTOL[np.dtype(dtype)] * 8 * np.max(np.abs(x))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_large_container_pieces_share_the_global_scaling_factor(dtype, tmp_path):
    """dctz_compress_large: block-aligned pieces, each a standard stream with the GLOBAL sf.
    With the piece length forced small: the concatenated bin indices equal the single-stream result, every
    piece decodes with the unmodified reference, and the dump tool walks the container."""
    lib = hostlib(False)
    lib.dctz_compress_large.restype = C.c_size_t
    lib.dctz_decompress_large.restype = C.c_size_t
    lib.dctz_large_bound.restype = C.c_size_t
    lib.dctz_compress_large.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_void_p, C.c_size_t]
    lib.dctz_decompress_large.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    lib.dctz_large_bound.argtypes = [C.c_size_t, C.c_int]
    lib.dctz_large_set_piece.argtypes = [C.c_size_t]
    x = np.concatenate([fields.small_cases(dtype)["smooth"] * 0.3 + 1.0, fields.small_cases(dtype)["tail32"]]).astype(dtype)  # max in the LAST piece
    n, eb, code = x.size, 1e-3, (1 if dtype == np.float64 else 0)
    piece = 64 * 300
    lib.dctz_large_set_piece(piece)
    try:
        cap = lib.dctz_large_bound(n, code)
        out = np.zeros(cap, np.uint8)
        size = lib.dctz_compress_large(x.ctypes.data, n, code, eb, out.ctypes.data, cap)
        assert 0 < size <= cap and bytes(out[:8]) == b"DCTZMS01"
        ns = int(np.frombuffer(out[20:24].tobytes(), "<u4")[0])
        assert ns == (n + piece - 1) // piece and int(np.frombuffer(out[8:16].tobytes(), "<u8")[0]) == n
        sizes = np.frombuffer(out[24:24 + 8 * ns].tobytes(), "<u8")
        whole, _, _ = our_compress(x, eb, False)
        hw, bins_w, dc_w, ac_w, _ = split_stream(whole, dtype, False)
        sf_w = np.frombuffer(hw["scaling_factor"].tobytes(), dtype=dtype)[0]
        off, bins, dcs, acs, recon_ref = 24 + 8 * ns, [], [], [], []
        for i in range(ns):
            s = out[off:off + int(sizes[i])]
            h, b, d, a, _ = split_stream(s, dtype, False)
            assert np.frombuffer(h["scaling_factor"].tobytes(), dtype=dtype)[0] == sf_w  # the GLOBAL scaling factor
            bins.append(b); dcs.append(d); acs.append(a)
            if reflib.have_ref():
                recon_ref.append(ref_decompress(s, int(h["num_elements"]), dtype, False))
            off += int(sizes[i])
        assert off == size
        assert np.array_equal(np.concatenate(bins), bins_w) and np.array_equal(np.concatenate(dcs), dc_w) and np.array_equal(np.concatenate(acs), ac_w)
        rec = np.zeros(n, dtype)
        assert lib.dctz_decompress_large(out.ctypes.data, size, rec.ctypes.data, n) == n
        want = our_decompress(whole, n, dtype, False)
        assert np.array_equal(rec, want)
        if recon_ref:
            tol = parity.RTOL[np.dtype(dtype)] * 8 * float(np.max(np.abs(x)))
            assert np.max(np.abs(np.concatenate(recon_ref).astype(np.float64) - rec.astype(np.float64))) <= tol
        f = tmp_path / "big.zms"
        out[:size].tofile(f)
        p = subprocess.run([os.path.join(ROOT, "dctz_b200", "bin", "dctz-dump"), str(f)], capture_output=True, text=True)
        assert p.returncode == 0 and f"streams={ns}" in p.stdout and p.stdout.count("SF=") == ns
    finally:
        lib.dctz_large_set_piece(1 << 30)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_quality_metrics_on_the_gpu(ctx, dtype):
    """dctz_gpu_quality = the reductions of calc_psnr (util.c:54-104)."""
    rng = np.random.default_rng(2)
    a = (rng.standard_normal(100003) * 3 + 1).astype(dtype)
    b = (a + rng.standard_normal(a.size).astype(dtype) * dtype(1e-3)).astype(dtype)
    q = ctx.quality(a, b)
    e = (a - b).astype(dtype).astype(np.float64)
    assert q["min"] == float(a.min()) and q["max"] == float(a.max()) and q["maxdiff"] == float(np.abs(e).max())
    assert abs(q["sumsq"] - float(np.sum(e * e))) <= 1e-9 * float(np.sum(e * e))
    lib = hostlib(False)
    va = reflib.TVar(1 if dtype == np.float64 else 0, 0.0, b"v", a.ctypes.data)
    vb = reflib.TVar(1 if dtype == np.float64 else 0, 0.0, b"v", b.ctypes.data)
    psnr = lib.calc_psnr(C.byref(va), C.byref(vb), C.c_int(a.size), C.c_double(1e-3))
    want = 20 * np.log10((float(a.max()) - float(a.min())) / np.sqrt(float(np.sum(e * e)) / a.size))
    assert abs(psnr - want) < 1e-9 * abs(want)


def test_concurrent_contexts_through_the_host_buffer_calls():
    """Three host threads with a context each push independent fields through dctz_gpu_compress_core / decompress_core on
    page-locked buffers at the same time (the per-device link gates put them in step; a dominant transfer keeps one piece
    queued while another call is active): every field must come out exactly as it does alone."""
    import threading

    import dctz_b200
    from dctz_b200 import binding

    n = 64 * 32 * 9000 + 64 * 3  # 147 MB of doubles: above the gates' threshold, streaming path
    rng = np.random.default_rng(21)
    t = np.arange(n, dtype=np.float64)
    fields_ = [3.0 + 2.0 * np.sin(t / (50.0 + 13 * k)) + 0.03 * rng.standard_normal(n) for k in range(3)]
    with dctz_b200.Context(0) as c0:
        alone = []
        for x in fields_:
            g = c0.compress_core(x, 1e-3)
            r = c0.decompress_core(g["bin_index"], g["dc"], g["ac"], n, np.float64, 1e-3, g["sf"])
            alone.append((g["bin_index"].copy(), g["dc"].copy(), g["ac"].copy(), r.copy()))
    ctxs = [dctz_b200.Context(0) for _ in fields_]
    pins = []
    for x in fields_:
        hx, hout = binding.PinnedArray((n,), np.float64), binding.PinnedArray((n,), np.float64)
        hb, hdc, hac = binding.PinnedArray((n,), np.uint8), binding.PinnedArray((n // 64,), np.float32), binding.PinnedArray((n,), np.float32)
        hx.array[:] = x
        pins.append((hx, hout, hb, hdc, hac))
    errs = []

    def work(k):
        try:
            hx, hout, hb, hdc, hac = pins[k]
            for _ in range(3):
                g = ctxs[k].compress_core(hx.array, 1e-3, out=dict(bin_index=hb.array, dc=hdc.array, ac_full=hac.array))
                ctxs[k].decompress_core(g["bin_index"], g["dc"], g["ac"], n, np.float64, 1e-3, g["sf"], out=hout.array)
                b, d, a, r = alone[k]
                assert np.array_equal(g["bin_index"], b) and np.array_equal(g["dc"], d) and np.array_equal(g["ac"], a)
                assert np.array_equal(hout.array, r)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(k,)) for k in range(3)]
    for t_ in th:
        t_.start()
    for t_ in th:
        t_.join(timeout=180)
    alive = any(t_.is_alive() for t_ in th)
    for c in ctxs:
        if not alive:
            c.close()
    for p in pins:
        for h in p:
            h.free()
    assert not alive, "concurrent host-buffer calls are stuck"
    assert not errs, errs
