"""CPU: pin the oracle (oracle/dctz_oracle.c) -- against genuine FFTW vectors, against the committed
outputs of the unmodified reference, against the reference itself when oracle/_ref is built, and
against an independent DCT (scipy/pocketfft)."""
import os
import sys

import numpy as np
import pytest

from dctz_b200 import fields
from tests import reflib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SIZES = make_golden.SIZES


@pytest.fixture(scope="module")
def fftw():
    return np.load(os.path.join(GOLD, "fftw_dct_ref.npz"))


@pytest.fixture(scope="module")
def refcases():
    return np.load(os.path.join(GOLD, "ref_cases.npz"))


@pytest.mark.parametrize("n", SIZES)
def test_oracle_dct_matches_fftw_double(fftw, n):
    x = np.linspace(0, n - 1, n)
    want = fftw[f"double_dct2_{n}"] / np.sqrt(2.0 * n)  # REDFT10 -> orthonormal DCT-II (dct.c:37-49, 100-102)
    want[0] /= np.sqrt(2.0)
    got = reflib.oracle_dct(x)
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
    back = reflib.oracle_dct(got, inverse=True)  # dct.c:115-205 inverts it
    assert np.max(np.abs(back - x)) <= 1e-12 * max(1.0, np.max(np.abs(x)))
    # REDFT01 (the unnormalised DCT-III) of the same ramp, through the oracle's inverse
    c = x.copy() * np.sqrt(2.0 * n) / (2.0 * n)
    c[0] *= np.sqrt(2.0)
    got3 = reflib.oracle_dct(c, inverse=True) * 2.0 * n / 2.0
    want3 = fftw[f"double_dct3_{n}"]
    assert np.max(np.abs(2 * got3 - want3)) <= 1e-11 * np.max(np.abs(want3))


@pytest.mark.parametrize("n", SIZES)
def test_oracle_dct_matches_fftw_single(fftw, n):
    x = np.linspace(0, n - 1, n).astype(np.float32)
    want = fftw[f"single_dct2_{n}"].astype(np.float64) / np.sqrt(2.0 * n)
    want[0] /= np.sqrt(2.0)
    got = reflib.oracle_dct(x).astype(np.float64)
    assert np.max(np.abs(got - want)) <= 1e-5 * np.max(np.abs(want))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_dct_matches_pocketfft_and_definition(dtype):
    from scipy.fft import dct, idct

    rng = np.random.default_rng(1)
    tol = 1e-13 if dtype == np.float64 else 2e-6
    for n in (64, 37, 17, 5, 1):
        x = rng.standard_normal(n).astype(dtype)
        got = reflib.oracle_dct(x).astype(np.float64)
        want = dct(x.astype(np.float64), type=2, norm="ortho")
        exact = reflib.oracle_dct_exact(x.astype(np.float64))
        scale = max(np.max(np.abs(want)), 1e-30)
        assert np.max(np.abs(got - want)) <= tol * scale * 8
        assert np.max(np.abs(exact - want)) <= 1e-13 * scale * 8
        inv = reflib.oracle_dct(want.astype(dtype), inverse=True).astype(np.float64)
        assert np.max(np.abs(inv - idct(want, type=2, norm="ortho"))) <= tol * 8 * max(1.0, np.max(np.abs(x)))


def test_conv_tbl_and_bin_centres():
    # dctz-comp-lib.c:27-43: ordinal 127 -> id 0, 128 -> 1, 126 -> 2 ... 0 -> 254, 254 -> 253; binning.c:19-22
    conv = [reflib.oracle().oracle_conv_tbl(t) for t in range(255)]
    assert conv[127] == 0 and conv[128] == 1 and conv[126] == 2 and conv[129] == 3 and conv[0] == 254 and conv[254] == 253
    assert sorted(conv) == list(range(255))
    eb = 1e-3
    centre = reflib.oracle_gen_bins(eb, np.float64)
    for t in range(255):  # bin t covers [(2t-255) eb, (2t-253) eb): its centre is (t-127) * 2eb
        assert abs(centre[conv[t]] - (t - 127) * 2 * eb) < 1e-15
    cf = reflib.oracle_gen_bins(eb, np.float32)
    assert cf.dtype == np.float32 and abs(float(cf[1]) - 2e-3) < 1e-9


def test_oracle_matches_committed_reference_outputs(refcases):
    """bit-for-bit against the fixtures produced by the unmodified reference (tools/make_golden.py)."""
    for name, (x, eb, qt) in make_golden.ref_case_inputs().items():
        o = reflib.oracle_compress(x, eb, qt)
        for key in ("bin_index", "dc", "ac", "scaled"):
            assert np.array_equal(o[key], refcases[f"{name}/{key}"]), (name, key)
        assert o["stat"]["sf"] == float(refcases[f"{name}/sf"][0]), name
        if qt:
            assert np.array_equal(o["qtable"], refcases[f"{name}/qtable"]), name
            assert np.array_equal(o["qtable_raw"], refcases[f"{name}/qtable_raw"]), name
        r = reflib.oracle_decompress(o["bin_index"], o["dc"], o["ac"], o["qtable"], x.size, eb, o["stat"]["sf"], qt, x.dtype)
        assert np.array_equal(r, refcases[f"{name}/recon"]), name
        assert o["n_edge"] == 0


@pytest.mark.skipif(not reflib.have_ref(), reason="oracle/_ref (the compiled reference) is not present")
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("qt", [False, True])
def test_oracle_matches_live_reference(dtype, qt):
    cases = fields.small_cases(dtype)
    for name in ("three_blocks_tail37", "tail32", "tail63_odd", "only_tail", "sf_one", "negative_only", "single_spike",
                 "heavy_outliers", "smooth", "all_equal"):
        x = cases[name]
        if x.size > 20000:
            x = x[:20000 + 32]
        r = reflib.ref_roundtrip(x, 1e-3, qt)
        o = reflib.oracle_compress(x, 1e-3, qt)
        assert np.array_equal(o["bin_index"], r["bin_index"]), name
        assert np.array_equal(o["dc"], r["dc"]), name
        assert np.array_equal(o["ac"], r["ac"]), name
        assert np.array_equal(o["coef"], r["coef"]), name
        assert np.array_equal(o["scaled"], r["scaled"]), name
        assert o["stat"]["sf"] == r["sf"], name
        if qt:
            assert np.array_equal(o["qtable"], r["qtable"]), name
        rec = reflib.oracle_decompress(o["bin_index"], o["dc"], o["ac"], o["qtable"], x.size, 1e-3, r["sf"], qt, dtype)
        assert np.array_equal(rec, r["recon"]), name


@pytest.mark.skipif(not reflib.have_ref(), reason="oracle/_ref (the compiled reference) is not present")
def test_reference_stream_layout():
    """header layout the host library reproduces (dctz.h:96-119): 56 bytes, sections, QT trailer."""
    import zlib

    x = fields.small_cases(np.float64)["tail32"][:6432]
    for qt in (False, True):
        r = reflib.ref_roundtrip(x, 1e-3, qt)
        h = r["header"]
        assert int(h["num_elements"]) == x.size and float(h["error_bound"]) == 1e-3 and int(h["datatype"]) == 1
        off = 56
        sizes = [int(h["bindex_sz_compressed"]), int(h["DC_sz_compressed"]), int(h["AC_exact_sz_compressed"])]
        secs = []
        for s in sizes:
            secs.append(zlib.decompress(r["stream"][off:off + s].tobytes()))
            off += s
        assert np.array_equal(np.frombuffer(secs[0], np.uint8), r["bin_index"])
        assert np.array_equal(np.frombuffer(secs[1], np.float32), r["dc"])
        assert np.array_equal(np.frombuffer(secs[2], np.float32), r["ac"])
        assert r["stream"].size == off + (512 if qt else 0)
        assert int(h["tot_AC_exact_count"]) == r["ac"].size


def test_error_bound_per_coefficient_in_oracle():
    """quirk 5 of SURVEY.md §8: the bound holds per non-outlier AC coefficient in the scaled domain."""
    x = fields.small_cases(np.float64)["smooth"]
    eb = 1e-3
    o = reflib.oracle_compress(x, eb, False)
    _, coef_r = reflib.oracle_decompress(o["bin_index"], o["dc"], o["ac"], None, x.size, eb, o["stat"]["sf"], False, np.float64,
                                        want_coef=True)
    inb = o["bin_index"] != 255
    assert np.max(np.abs(o["coef"][inb] - coef_r[inb])) <= eb * (1 + 1e-12)
