"""TEST INFRASTRUCTURE: ctypes access to oracle/liboracle.so (our CPU restatement) and to
oracle/_ref/*.so (the unmodified reference compiled against the FFTW stand-in)."""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


class OracleStat(C.Structure):
    _fields_ = [("max", C.c_double), ("min", C.c_double), ("sum", C.c_double), ("mean", C.c_double), ("sf", C.c_double)]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            raise RuntimeError("oracle/liboracle.so missing: run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C oracle`")
        _oracle = C.CDLL(ORACLE_SO)
        _oracle.oracle_conv_tbl.restype = C.c_ubyte
    return _oracle


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libdctz_ref_ec.so"))


def _T(dtype):
    dtype = np.dtype(dtype)
    assert dtype in (np.float64, np.float32)
    return ("d", np.float64) if dtype == np.float64 else ("f", np.float32)


def oracle_dct(x, inverse=False):
    sfx, dt = _T(x.dtype)
    x = np.ascontiguousarray(x)
    out = np.empty_like(x)
    fn = getattr(oracle(), f"oracle_{'idct' if inverse else 'dct'}_{sfx}")
    fn(_ptr(x), _ptr(out), C.c_int(x.size))
    return out


def oracle_dct_exact(x, inverse=False):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    fn = oracle().oracle_idct_exact if inverse else oracle().oracle_dct_exact
    fn(_ptr(x), _ptr(out), C.c_int(x.size))
    return out


def oracle_stat(x):
    sfx, dt = _T(x.dtype)
    st = OracleStat()
    getattr(oracle(), f"oracle_calc_stat_{sfx}")(_ptr(np.ascontiguousarray(x)), C.c_long(x.size), C.byref(st))
    return dict(max=st.max, min=st.min, sum=st.sum, mean=st.mean, sf=st.sf)


def oracle_gen_bins(eb, dtype):
    sfx, dt = _T(dtype)
    c = np.empty(255, dtype=dt)
    if sfx == "d":
        oracle().oracle_gen_bins_d(_ptr(c), 255, C.c_double(eb))
    else:
        oracle().oracle_gen_bins_f(_ptr(c), 255, C.c_float(np.float32(eb)))
    return c


def oracle_compress(x, eb, qt, want_coef=True):
    """Returns dict with scaled (the in-place mutated input), bin_index, dc, ac, qtable_raw, qtable, stat, coef, n_edge."""
    sfx, dt = _T(x.dtype)
    n = x.size
    nblk = (n + 63) // 64
    buf = np.array(x, dtype=dt, copy=True)
    bins = np.empty(n, dtype=np.uint8)
    dc = np.empty(nblk, dtype=np.float32)
    ac = np.empty(max(n, 1), dtype=np.float32)
    n_out = C.c_uint(0)
    qraw = np.zeros(64, dtype=dt)
    qtab = np.zeros(64, dtype=dt)
    coef = np.zeros(n, dtype=dt) if want_coef else None
    st = OracleStat()
    edge = C.c_ulong(0)
    fn = getattr(oracle(), f"oracle_compress_core_{sfx}")
    fn(_ptr(buf), C.c_long(n), C.c_double(eb), C.c_int(int(qt)), _ptr(bins), _ptr(dc), _ptr(ac), C.byref(n_out),
       _ptr(qraw), _ptr(qtab), C.byref(st), _ptr(coef), C.byref(edge))
    return dict(scaled=buf, bin_index=bins, dc=dc, ac=ac[: n_out.value].copy(), qtable_raw=qraw, qtable=qtab,
                stat=dict(max=st.max, min=st.min, sum=st.sum, mean=st.mean, sf=st.sf), coef=coef, n_edge=edge.value)


def oracle_decompress(bins, dc, ac, qtable, n, eb, sf, qt, dtype, want_coef=False):
    sfx, dt = _T(dtype)
    out = np.empty(n, dtype=dt)
    coef = np.zeros(n, dtype=dt) if want_coef else None
    ac = np.ascontiguousarray(ac, dtype=np.float32)
    if ac.size == 0:
        ac = np.zeros(1, dtype=np.float32)
    q = None if qtable is None else np.ascontiguousarray(qtable, dtype=dt)
    fn = getattr(oracle(), f"oracle_decompress_core_{sfx}")
    sfarg = C.c_double(sf) if sfx == "d" else C.c_float(np.float32(sf))
    fn(_ptr(np.ascontiguousarray(bins)), _ptr(np.ascontiguousarray(dc, dtype=np.float32)), _ptr(ac), _ptr(q), C.c_long(n),
       C.c_double(eb), sfarg, C.c_int(int(qt)), _ptr(out), _ptr(coef))
    return (out, coef) if want_coef else out


# --------------------------------------------------------------------------------------------
# the unmodified reference (oracle/_ref)
# --------------------------------------------------------------------------------------------


class TVar(C.Structure):  # dctz.h:49-59
    _fields_ = [("datatype", C.c_int), ("err_bound", C.c_double), ("var_name", C.c_char_p), ("buf", C.c_void_p)]


HEADER_DTYPE = np.dtype([  # dctz.h:96-119; 56 bytes in both modes
    ("datatype", "<i4"), ("num_elements", "<u4"), ("error_bound", "<f8"), ("tot_AC_exact_count", "<u4"), ("_pad0", "<u4"),
    ("scaling_factor", "V8"), ("mean", "V8"), ("bindex_sz_compressed", "<u4"), ("DC_sz_compressed", "<u4"),
    ("AC_exact_sz_compressed", "<u4"), ("bindex_count", "<u4")])

_ref_libs = {}


def ref_lib(qt):
    key = "qt" if qt else "ec"
    if key not in _ref_libs:
        _ref_libs[key] = C.CDLL(os.path.join(REF_DIR, f"libdctz_ref_{key}.so"))
    return _ref_libs[key]


@contextlib.contextmanager
def _in_tmpdir():
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            yield d
        finally:
            os.chdir(old)


def ref_roundtrip(x, eb, qt):
    """Run the reference's dctz_compress + dctz_decompress in-process (dctz-test.c:181,250) and collect
    the stream, its debug dumps (dctz-comp-lib.c:422-448, 583-595) and the reconstruction."""
    sfx, dt = _T(x.dtype)
    lib = ref_lib(qt)
    n = x.size
    nblk = (n + 63) // 64
    code = 1 if sfx == "d" else 0  # dctz.h:44-47 FLOAT = 0, DOUBLE = 1
    buf = np.array(x, dtype=dt, copy=True)
    zbuf = np.zeros(2 * n * dt().itemsize + 4096, dtype=np.uint8)
    rbuf = np.zeros(n, dtype=dt)
    var = TVar(code, eb, b"v", buf.ctypes.data)
    var_z = TVar(code, eb, b"v", zbuf.ctypes.data)
    var_r = TVar(code, eb, b"v", rbuf.ctypes.data)
    out_size = C.c_size_t(0)
    with _in_tmpdir():
        lib.dctz_compress(C.byref(var), C.c_int(n), C.byref(out_size), C.byref(var_z), C.c_double(eb))
        dumps = dict(
            bin_index=np.fromfile("bin_index.bin", dtype=np.uint8),
            ac=np.fromfile("AC_exact.bin", dtype=np.float32),
            coef=np.fromfile("dct_result.bin", dtype=dt),
            dc=np.fromfile("DC.bin", dtype=np.float32),
        )
        if qt:
            dumps["qtable_raw"] = np.fromfile("qtable.bin", dtype=dt)
        lib.dctz_decompress(C.byref(var_z), C.byref(var_r))
    stream = zbuf[: out_size.value].copy()
    hdr = np.frombuffer(stream[:56].tobytes(), dtype=HEADER_DTYPE)[0]
    sf = np.frombuffer(hdr["scaling_factor"].tobytes(), dtype=dt)[0]
    mean = np.frombuffer(hdr["mean"].tobytes(), dtype=dt)[0]
    assert dumps["bin_index"].size == n and dumps["dc"].size == nblk
    res = dict(scaled=buf, stream=stream, header=hdr, sf=float(sf), mean=float(mean), recon=rbuf, **dumps)
    if qt:
        res["qtable"] = np.frombuffer(stream[-64 * dt().itemsize:].tobytes(), dtype=dt).copy()
    return res


def ref_dct(x, inverse=False):
    """dct_init/dct_fftw/dct_finish and ifft_idct/idct_finish of the reference (dct.h:17-27)."""
    sfx, dt = _T(x.dtype)
    lib = ref_lib(False)
    x = np.ascontiguousarray(x)
    out = np.zeros_like(x)
    f = "_f" if sfx == "f" else ""
    n = C.c_int(x.size)
    if not inverse:
        getattr(lib, "dct_init" + f)(n)
        getattr(lib, "dct_fftw" + f)(_ptr(x), _ptr(out), n, C.c_int(1))
        getattr(lib, "dct_finish" + f)()
    else:
        getattr(lib, "ifft_idct" + f)(n, _ptr(x), _ptr(out))
        getattr(lib, "idct_finish" + f)()
    return out
