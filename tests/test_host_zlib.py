"""CPU: the host library's zlib stage.  Each stream section must be ONE valid zlib
stream that any inflate() decodes -- also when it was deflated chunk-parallel on all host cores --, and the
single-threaded path must reproduce the reference's deflate parameters byte for byte."""
import ctypes as C
import os
import subprocess
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    lib = C.CDLL(os.path.join(ROOT, "dctz_b200", "libdctz_ec.so"))
    lib.dctz_host_deflate.restype = C.c_size_t
    lib.dctz_host_deflate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    return lib


def _deflate(data: bytes) -> bytes:
    lib = _lib()
    cap = len(data) + len(data) // 8 + 4096
    out = C.create_string_buffer(cap)
    src = C.create_string_buffer(data, len(data)) if data else C.create_string_buffer(1)
    n = lib.dctz_host_deflate(src, len(data), out, cap)
    assert n > 0
    return out.raw[:n]


def _bin_index_like(n, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.choice(np.array([0, 1, 2, 3, 4, 255], np.uint8), size=n, p=[0.55, 0.15, 0.12, 0.08, 0.05, 0.05])
    a[::64] = 255
    return a.tobytes()


@pytest.mark.parametrize("n", [0, 1, 1000, (1 << 21), (1 << 21) + 1, 5 * (1 << 20) + 12345, 16 << 20])
def test_section_is_one_valid_zlib_stream(n):
    data = _bin_index_like(n)
    z = _deflate(data)
    assert zlib.decompress(z) == data  # standard inflate, as in dctz-decomp-lib.c:244-322
    d = zlib.decompressobj()
    assert d.decompress(z) == data and d.eof and d.unused_data == b""  # exactly one stream, nothing trailing


def test_small_sections_match_the_reference_parameters():
    data = _bin_index_like(300000, seed=3)  # below the parallel threshold: deflateInit2(default, 15, 8) + Z_FINISH
    assert _deflate(data) == zlib.compress(data, -1)


def test_parallel_ratio_is_close_to_serial_and_threads_env_is_honoured():
    data = _bin_index_like(24 << 20, seed=5)
    par = _deflate(data)
    ser = zlib.compress(data, -1)
    assert len(par) <= len(ser) * 1.01 + 64  # a sync-flush marker per 1 MiB chunk, dictionary carried over
    code = ("import sys; sys.path.insert(0, %r); from tests.test_host_zlib import _deflate, _bin_index_like; import zlib; "
            "d = _bin_index_like(5 << 20, seed=7); assert _deflate(d) == zlib.compress(d, -1); print('same')" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DCTZ_ZLIB_THREADS="1"), capture_output=True, text=True)
    assert p.returncode == 0 and "same" in p.stdout, p.stderr


@pytest.mark.parametrize("n,piece", [(0, 0), (1000, 300), ((1 << 21) + 1, 1 << 20), (5 * (1 << 20) + 12345, 700001), (16 << 20, 16 << 20),
                                     (24 << 20, (16 << 20))])
def test_streamed_deflate_is_byte_identical(n, piece):
    """dctz_compress deflates the sections while they arrive from the GPU in pieces (zpipe): the stream must be the one
    deflate_sections produces from the complete array, whatever the piece size."""
    lib = _lib()
    lib.dctz_host_deflate_streamed.restype = C.c_size_t
    lib.dctz_host_deflate_streamed.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t]
    data = _bin_index_like(n, seed=11)
    cap = n + n // 8 + 4096
    out = C.create_string_buffer(cap)
    src = C.create_string_buffer(data, len(data)) if data else C.create_string_buffer(1)
    m = lib.dctz_host_deflate_streamed(src, n, out, cap, piece)
    assert m > 0 and out.raw[:m] == _deflate(data)
    assert zlib.decompress(out.raw[:m]) == data


@pytest.mark.parametrize("section", [1, 2])
@pytest.mark.parametrize("n,piece", [(0, 0), (300000, 0), ((1 << 21) + 4, 0), (9 << 20, 0), (9 << 20, 1 << 20), ((8 << 20) + 12, 700000)])
def test_float_sections_use_small_chunks_and_stay_one_stream(section, n, piece):
    """DC / AC_exact (float data, slow to deflate) are cut into 128 KiB chunks: still ONE valid zlib stream, the same bytes
    whether the section is deflated whole or while it arrives, and close to the serial size."""
    lib = _lib()
    lib.dctz_host_deflate_section.restype = C.c_size_t
    lib.dctz_host_deflate_section.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_size_t]
    rng = np.random.default_rng(17)
    data = (np.cumsum(rng.standard_normal(n // 4)) * 0.01).astype(np.float32).tobytes()
    n = len(data)
    cap = n + n // 8 + 4096
    src = C.create_string_buffer(data, n) if n else C.create_string_buffer(1)
    out = C.create_string_buffer(cap)
    m = lib.dctz_host_deflate_section(src, n, out, cap, section, piece)
    assert m > 0
    z = out.raw[:m]
    d = zlib.decompressobj()
    assert d.decompress(z) == data and d.eof and d.unused_data == b""
    whole = C.create_string_buffer(cap)
    mw = lib.dctz_host_deflate_section(src, n, whole, cap, section, 0)
    assert whole.raw[:mw] == z
    assert m <= len(zlib.compress(data, -1)) * 1.01 + 64
