"""CPU: the generated straight-line DCT-64 (tools/gen_dct64.py -> dctz_b200/csrc/dct64_gen.cuh)."""
import os
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_dct64 as g  # noqa: E402


@pytest.fixture(scope="module")
def progs():
    return g.build_forward(64), g.build_inverse(64)


def test_flow_graph_is_the_orthonormal_dct(progs):
    from scipy.fft import dct

    f, i = progs
    x = np.random.default_rng(0).standard_normal((500, 64))
    want = dct(x, type=2, norm="ortho", axis=-1)
    got = g.evaluate(f, x, np.float64)
    assert np.max(np.abs(got - want)) <= 2e-15 * np.max(np.abs(want))
    back = g.evaluate(i, want, np.float64)
    assert np.max(np.abs(back - x)) <= 4e-15 * np.max(np.abs(x))


def test_float_accuracy_inside_tolerance(progs):
    from scipy.fft import dct

    f, i = progs
    x = np.random.default_rng(1).standard_normal((500, 64))
    want = dct(x, type=2, norm="ortho", axis=-1)
    got = g.evaluate(f, x.astype(np.float32), np.float32)
    scale = np.max(np.abs(want), axis=-1, keepdims=True)
    assert np.max(np.abs(got - want) / scale) <= 1e-6  # 1e-5 is the stated tolerance
    back = g.evaluate(i, want.astype(np.float32), np.float32)
    assert np.max(np.abs(back - x)) <= 4e-6


def test_operation_count_and_symmetry(progs):
    f, i = progs
    cf, ci = g.op_counts(f), g.op_counts(i)
    assert cf["total"] == 592 and ci["total"] == 592  # documented in DESIGN.md (9.25 flop-instructions / element)
    assert cf["fma"] == ci["fma"] and cf["mul"] == ci["mul"]


def test_committed_header_is_up_to_date():
    path = os.path.join(ROOT, "dctz_b200", "csrc", "dct64_gen.cuh")
    with tempfile.TemporaryDirectory() as d:
        tmp = os.path.join(d, "dct64_gen.cuh")
        old = sys.argv
        sys.argv = ["gen_dct64.py", tmp]
        try:
            g.main()
        finally:
            sys.argv = old
        assert open(tmp).read() == open(path).read(), "run tools/gen_dct64.py and commit the result"


def test_constant_table_agrees_with_the_literals():
    """The double policy reads entry i of dct64_kd, the float policy the literal beside it (A::cst(i, literal)): every use
    of an index must name the literal the table holds at that index, and every entry must be used."""
    import re

    path = os.path.join(ROOT, "dctz_b200", "csrc", "dct64_gen.cuh")
    txt = open(path).read()
    nk = int(re.search(r"constexpr int DCT64_NK = (\d+);", txt).group(1))
    body = re.search(r"dct64_kd_host\[DCT64_NK\] = \{(.*?)\};", txt, re.S).group(1)
    table = [v.strip() for v in body.split(",") if v.strip()]
    assert len(table) == nk == len(set(table))
    uses = re.findall(r"A::cst\((\d+), (-?[0-9.eE+-]+)\)", txt)
    assert len(uses) == 2 * (136 + 114)  # mul + fma of the forward and of the inverse flow graph
    assert all(table[int(i)] == lit for i, lit in uses)
    assert {int(i) for i, _ in uses} == set(range(nk))
