"""TEST INFRASTRUCTURE: parity checks of the CUDA path (through the C-ABI) against the CPU oracle.

Tolerances are the ones BASELINE.json's north_star states: DCT coefficients within 1e-12 relative
(double) / 1e-5 (float) -- relative to the largest coefficient magnitude of the block --, bin indices
and outlier sets bit-exact except at quantisation-boundary ties, whose count is reported.  A "tie" is a
coefficient whose oracle value lies within the coefficient tolerance of a bin boundary (or of the
outlier range limit), so that two correct DCT implementations may legitimately round it to either side.
"""
from __future__ import annotations

import numpy as np

from . import reflib

RTOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


def quant_consts(eb, dtype):
    """dctz-comp-lib.c:271-281."""
    dtype = np.dtype(dtype)
    bw, rmin, rmax = eb * 2.0, -255 * eb, 255 * eb
    if dtype == np.float32:
        bw, rmin, rmax = np.float32(bw), np.float32(rmin), np.float32(rmax)
    return float(bw), float(rmin), float(rmax)


def block_max(coef):
    """max |coef| of the 64-block each element belongs to (same shape as coef)."""
    n = coef.size
    nblk = (n + 63) // 64
    pad = np.zeros(nblk * 64, dtype=np.float64)
    pad[:n] = np.abs(coef)
    m = pad.reshape(nblk, 64).max(axis=1)
    return np.repeat(m, 64)[:n]


def boundary_distance(coef, eb, dtype):
    """distance of each coefficient to the nearest bin boundary rmin + t*bw, t = 0..255."""
    bw, rmin, rmax = quant_consts(eb, dtype)
    c = coef.astype(np.float64)
    v = (c - rmin) / bw
    t = np.clip(np.rint(v), 0, 255)
    return np.abs(v - t) * bw


def ulp32(a):
    a = np.abs(a.astype(np.float32))
    return (np.nextafter(a, np.float32(np.inf)) - a).astype(np.float64)


def compare_compress(gpu, orc, x, eb, qt):
    """gpu: result of Context.compress_core; orc: reflib.oracle_compress (or reference dumps with the
    same keys).  Asserts parity and returns a report."""
    dtype = np.dtype(x.dtype)
    rtol = RTOL[dtype]
    n = x.size
    coef = orc["coef"].astype(np.float64)
    bmax = block_max(coef)
    ctol = rtol * np.maximum(bmax, 1e-300)
    if dtype == np.float32:
        # the reference evaluates (c - range_min)/bin_width in float: one ulp of a value up to 255 is the
        # resolution of that expression itself, so a boundary closer than that is a tie as well
        ctol = ctol + 255.0 * 2.0 ** -23 * quant_consts(eb, dtype)[0]
    rep = {}

    # statistics (util.c:12-44): max/min are exact selections, sf must be bit-identical
    st = orc["stat"]
    info = gpu["info"]
    assert info["max_abs"] == st["max"], (info["max_abs"], st["max"])
    assert info["min_abs"] == st["min"], (info["min_abs"], st["min"])
    assert info["sf"] == st["sf"], (info["sf"], st["sf"])
    sum_tol = (1e-9 if dtype == np.float64 else 2e-3) * max(1.0, float(np.sum(np.abs(x.astype(np.float64)))))
    assert abs(info["sum"] - st["sum"]) <= sum_tol, (info["sum"], st["sum"])

    # bin indices
    gb, ob = gpu["bin_index"], orc["bin_index"]
    assert gb.shape == ob.shape
    diff = np.nonzero(gb != ob)[0]
    rep["bin_mismatch"] = int(diff.size)
    if diff.size:
        dist = boundary_distance(coef[diff], eb, dtype)
        not_tie = diff[dist > ctol[diff]]
        assert not_tie.size == 0, f"{not_tie.size} bin indices differ away from a boundary, first at {not_tie[:5]}: " \
                                  f"gpu {gb[not_tie[:5]]} oracle {ob[not_tie[:5]]} coef {coef[not_tie[:5]]}"
        # a tie may only move to the neighbouring bin (or in/out of the outlier range)
    rep["ties"] = int(diff.size)
    rep["tie_fraction"] = diff.size / max(n, 1)

    # DC (dctz-comp-lib.c:351): float of the block's first coefficient
    gdc, odc = gpu["dc"].astype(np.float64), orc["dc"].astype(np.float64)
    dc_tol = ctol[::64] + ulp32(orc["dc"])
    bad = np.nonzero(np.abs(gdc - odc) > dc_tol)[0]
    assert bad.size == 0, f"DC differs in {bad.size} blocks, first {bad[:5]}: {gdc[bad[:5]]} vs {odc[bad[:5]]}"

    # outliers: same set (up to ties), same order, same float values
    pos = np.arange(n) % 64
    gmask = (gb == 255) & (pos != 0)
    omask = (ob == 255) & (pos != 0)
    rep["n_outliers"] = int(info["n_outliers"])
    assert gpu["ac"].size == info["n_outliers"]
    if not qt or info["n_qt_dropped"] == 0:
        assert int(gmask.sum()) == info["n_outliers"], (int(gmask.sum()), info["n_outliers"])
    gval = np.full(n, np.nan)
    oval = np.full(n, np.nan)
    if not qt or info["n_qt_dropped"] == 0:
        gval[gmask] = gpu["ac"]
    oval[omask] = orc["ac"] if orc["ac"].size == int(omask.sum()) else np.nan
    both = gmask & omask & ~np.isnan(gval) & ~np.isnan(oval)
    if qt:
        # rescaled value = range +- (c/qtable[j]) * 10 eb: sensitivity to c is 10 eb / qtable[j] <= 10 eb
        ac_tol = 10 * eb * ctol[both] + 2 * ulp32(oval[both].astype(np.float32))
    else:
        ac_tol = ctol[both] + ulp32(oval[both].astype(np.float32))
    bad = np.nonzero(np.abs(gval[both] - oval[both]) > ac_tol)[0]
    assert bad.size == 0, f"{bad.size} outlier values differ, e.g. {gval[both][bad[:5]]} vs {oval[both][bad[:5]]}"
    rep["outlier_set_diff"] = int((gmask != omask).sum())
    assert rep["outlier_set_diff"] <= rep["ties"]

    if qt:
        # qtable (dctz-comp-lib.c:371-372, 450-461): a max over existing values -> exact unless a tie changed the set
        gq, oq = gpu["qtable"].astype(np.float64), orc["qtable"].astype(np.float64)
        qtol = rtol * max(1.0, float(np.max(np.abs(coef)))) * 4
        assert np.all(np.abs(gq - oq) <= qtol), (gq, oq)
        rep["qtable_exact"] = bool(np.array_equal(gpu["qtable"][1:], orc["qtable"][1:]))
    return rep


def check_compress(ctx, x, eb, qt):
    gpu = ctx.compress_core(x, eb, qt=qt, want_scaled=True)
    orc = reflib.oracle_compress(x, eb, qt)
    rep = compare_compress(gpu, orc, x, eb, qt)
    # the in-place scaling the reference leaves in the caller's buffer: IEEE division, bit-exact
    assert np.array_equal(gpu["scaled"], orc["scaled"]), "x/sf is not bit-identical to the reference's division"
    return rep


def check_decompress(ctx, x, eb, qt):
    """Feed the ORACLE's compressed arrays to the GPU decompressor; compare with the oracle's own
    reconstruction (dctz-decomp-lib.c:358-511)."""
    dtype = np.dtype(x.dtype)
    orc = reflib.oracle_compress(x, eb, qt, want_coef=False)
    sf = orc["stat"]["sf"]
    want = reflib.oracle_decompress(orc["bin_index"], orc["dc"], orc["ac"], orc["qtable"], x.size, eb, sf, qt, dtype)
    got = ctx.decompress_core(orc["bin_index"], orc["dc"], orc["ac"], x.size, dtype, eb, sf, qt=qt, qtable=orc["qtable"])
    scale = float(np.max(np.abs(want))) if want.size else 1.0
    diff = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)))) if want.size else 0.0
    assert diff <= RTOL[dtype] * max(scale, 1e-300) * 8, (diff, scale)
    return dict(max_diff=diff, scale=scale)
