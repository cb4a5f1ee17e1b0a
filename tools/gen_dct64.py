#!/usr/bin/env python3
"""Generate straight-line, register-resident 64-point orthonormal DCT-II / DCT-III code.

What the reference computes per 64-element block (dct.c:55-103 forward, dct.c:115-205 inverse)
is the *orthonormal* DCT-II and its inverse (DCT-III); it gets there through FFTW (Makhoul's
reorder + complex FFT + twiddle).  On the GPU one thread owns one block in registers, so we want
a flow graph with as few FP operations as possible, no data-dependent indexing, and orthogonal
stages only (numerical error ~ log2(N)*eps, needed for the 1e-5 float tolerance).

Factorisation (derived in DESIGN.md, all stages are butterflies or plane rotations):

  DCT-II_N(x):  u[n] = x[n] + x[N-1-n], v[n] = x[n] - x[N-1-n]            (n < N/2)
                X[2m]   = DCT-II_{N/2}(u)[m]
                X[2m+1] = DCT-IV_{N/2}(v)[m]
  DCT-IV_M(v):  H = M/2, alpha_n = (2n+1)pi/(4M)
                a[n] =  v[n] cos(alpha_n) + v[M-1-n] sin(alpha_n)
                b[n] = -v[n] sin(alpha_n) + v[M-1-n] cos(alpha_n)            (n < H)
                Ca = DCT-II_H(a),  Sb[p] = DCT-II_H((-1)^n b[n])[H-1-p]   (= DST-II_H(b))
                Y[0] = Ca[0];  Y[2p] = Ca[p] + Sb[p-1];  Y[2p+1] = Ca[p+1] - Sb[p];  Y[M-1] = -Sb[H-1]

The orthonormal scale sqrt(2/N) (and the extra 1/sqrt2 of the DC term) is folded into rotation
constants, so it costs a single extra multiply.  The inverse is generated as the exact transpose
of the forward flow graph (the transform is orthogonal), so both directions have the same
operation count.

Outputs:
  dctz_b200/csrc/dct64_gen.cuh   (device code, templated on an arithmetic policy)
Also usable as a module: build_forward(n) / build_inverse(n) return op lists that
`evaluate()` can run under numpy in float64 or float32 (used by tests/test_dct_codegen.py).
"""
from __future__ import annotations

import math
import os
import sys
from dataclasses import dataclass, field

try:
    import mpmath

    mpmath.mp.prec = 200

    def _cos(num, den):  # cos(pi*num/den)
        return mpmath.cos(mpmath.pi * num / den)

    def _sin(num, den):
        return mpmath.sin(mpmath.pi * num / den)

    def _sqrt(x):
        return mpmath.sqrt(x)

    def _mk(x):
        return mpmath.mpf(x)

except ImportError:  # pragma: no cover - numpy longdouble fallback
    import numpy as _np

    def _cos(num, den):
        return _np.cos(_np.longdouble(_np.pi) * num / den)

    def _sin(num, den):
        return _np.sin(_np.longdouble(_np.pi) * num / den)

    def _sqrt(x):
        return _np.sqrt(_np.longdouble(x))

    def _mk(x):
        return _np.longdouble(x)


# ----------------------------------------------------------------------------------------------
# A tiny SSA builder for linear flow graphs.  A "signed value" is (id, sign); negations are never
# emitted, they are folded into the consumer (sub instead of add, negated constant, ...).
# ----------------------------------------------------------------------------------------------


@dataclass
class Prog:
    n_in: int
    ops: list = field(default_factory=list)  # (kind, dst, a, b, k)
    n_val: int = 0
    outputs: list = field(default_factory=list)  # [(id, sign)]

    def __post_init__(self):
        self.n_val = self.n_in

    def _new(self, kind, a=None, b=None, k=None):
        dst = self.n_val
        self.n_val += 1
        self.ops.append((kind, dst, a, b, k))
        return dst

    # value algebra -------------------------------------------------------------------------
    def add(self, x, y):
        (a, sa), (b, sb) = x, y
        if sa > 0 and sb > 0:
            return (self._new("add", a, b), 1)
        if sa > 0 and sb < 0:
            return (self._new("sub", a, b), 1)
        if sa < 0 and sb > 0:
            return (self._new("sub", b, a), 1)
        return (self._new("add", a, b), -1)

    def sub(self, x, y):
        return self.add(x, (y[0], -y[1]))

    def mul(self, x, k):
        (a, sa) = x
        return (self._new("mul", a, None, k * sa), 1)

    def fma(self, x, k, y):
        """x*k + y"""
        (a, sa), (b, sb) = x, y
        k = k * sa
        if sb > 0:
            return (self._new("fma", a, b, k), 1)  # a*k + b
        return (self._new("fms", a, b, k), 1)  # a*k - b


def _dct2(p: Prog, xs, scale, dc):
    """Unnormalised DCT-II of the signed values xs, every output multiplied by `scale`
    (output 0 additionally by 1/sqrt2 when dc is True)."""
    n = len(xs)
    if n == 1:
        s = scale * (1 / _sqrt(2) if dc else 1)
        return [xs[0] if s == 1 else p.mul(xs[0], s)]
    if n == 2:
        s0 = scale * (1 / _sqrt(2) if dc else 1)
        e = p.add(xs[0], xs[1])
        o = p.sub(xs[0], xs[1])
        if s0 != 1:
            e = p.mul(e, s0)
        o = p.mul(o, scale * _cos(1, 4))
        return [e, o]
    h = n // 2
    u = [p.add(xs[i], xs[n - 1 - i]) for i in range(h)]
    v = [p.sub(xs[i], xs[n - 1 - i]) for i in range(h)]
    ev = _dct2(p, u, scale, dc)
    od = _dct4(p, v, scale)
    out = [None] * n
    out[0::2] = ev
    out[1::2] = od
    return out


def _dct4(p: Prog, vs, scale):
    """Unnormalised DCT-IV: Y[m] = sum v[n] cos((2n+1)(2m+1)pi/(4M)), times `scale`."""
    m = len(vs)
    if m == 1:
        return [p.mul(vs[0], scale * _cos(1, 4))]
    h = m // 2
    a, b = [], []
    for i in range(h):
        c = _cos(2 * i + 1, 4 * m) * scale
        s = _sin(2 * i + 1, 4 * m) * scale
        x, y = vs[i], vs[m - 1 - i]
        # a =  x c + y s ; b = -x s + y c   (2 mul + 2 fma)
        a.append(p.fma(x, c, p.mul(y, s)))
        bi = p.fma(y, c, p.mul(x, -s))
        if i & 1:
            bi = (bi[0], -bi[1])  # (-1)^n b[n]
        b.append(bi)
    ca = _dct2(p, a, _mk(1), False)
    cb = _dct2(p, b, _mk(1), False)
    sb = cb[::-1]  # Sb[p] = Cb[H-1-p]
    y = [None] * m
    y[0] = ca[0]
    for q in range(1, h):
        y[2 * q] = p.add(ca[q], sb[q - 1])
    for q in range(0, h - 1):
        y[2 * q + 1] = p.sub(ca[q + 1], sb[q])
    y[m - 1] = (sb[h - 1][0], -sb[h - 1][1])
    return y


def build_forward(n=64) -> Prog:
    p = Prog(n)
    xs = [(i, 1) for i in range(n)]
    scale = _sqrt(_mk(2) / n)
    outs = _dct2(p, xs, scale, True)
    p.outputs = outs
    return p


def build_inverse(n=64) -> Prog:
    """Transpose the forward flow graph: reverse-mode accumulation over the linear program."""
    f = build_forward(n)
    p = Prog(n)
    adj = {}  # forward value id -> signed value in p

    def contribute(fid, val):
        if fid in adj:
            adj[fid] = p.add(adj[fid], val)
        else:
            adj[fid] = val

    def contribute_scaled(fid, val, k):
        if fid in adj:
            adj[fid] = p.fma(val, k, adj[fid])
        else:
            adj[fid] = p.mul(val, k)

    for k, (fid, sgn) in enumerate(f.outputs):
        contribute(fid, (k, sgn))
    for kind, dst, a, b, k in reversed(f.ops):
        d = adj.pop(dst)
        if kind == "add":
            contribute(a, d)
            contribute(b, d)
        elif kind == "sub":
            contribute(a, d)
            contribute(b, (d[0], -d[1]))
        elif kind == "mul":
            contribute_scaled(a, d, k)
        elif kind == "fma":  # dst = a*k + b
            contribute(b, d)
            contribute_scaled(a, d, k)
        elif kind == "fms":  # dst = a*k - b
            contribute(b, (d[0], -d[1]))
            contribute_scaled(a, d, k)
        else:  # pragma: no cover
            raise AssertionError(kind)
    p.outputs = [adj[i] for i in range(n)]
    return p


# ----------------------------------------------------------------------------------------------
# numpy evaluation (for tests) and statistics
# ----------------------------------------------------------------------------------------------


def evaluate(p: Prog, x, dtype):
    """x: array [..., n_in]; every op is rounded to `dtype` (fma emulated in float64/longdouble)."""
    import numpy as np

    wide = np.longdouble if dtype == np.float64 else np.float64
    vals = [None] * p.n_val
    for i in range(p.n_in):
        vals[i] = x[..., i].astype(dtype)
    for kind, dst, a, b, k in p.ops:
        kk = dtype(float(k)) if k is not None else None
        if kind == "add":
            vals[dst] = vals[a] + vals[b]
        elif kind == "sub":
            vals[dst] = vals[a] - vals[b]
        elif kind == "mul":
            vals[dst] = vals[a] * kk
        elif kind == "fma":
            vals[dst] = (vals[a].astype(wide) * wide(kk) + vals[b].astype(wide)).astype(dtype)
        elif kind == "fms":
            vals[dst] = (vals[a].astype(wide) * wide(kk) - vals[b].astype(wide)).astype(dtype)
    outs = [vals[i] if s > 0 else -vals[i] for (i, s) in p.outputs]
    return np.stack(outs, axis=-1)


def op_counts(p: Prog):
    c = {}
    for kind, *_ in p.ops:
        c[kind] = c.get(kind, 0) + 1
    c["total"] = len(p.ops)
    return c


# ----------------------------------------------------------------------------------------------
# CUDA emission
# ----------------------------------------------------------------------------------------------


def _lit(k):
    try:
        return mpmath.nstr(k, 21, strip_zeros=False)
    except NameError:  # pragma: no cover
        return repr(float(k))


def emit_function(p: Prog, name: str, table: dict) -> str:
    """In-place transform of `V (&x)[N]`.  A is the arithmetic policy (common.cuh):
    A::add(a,b) A::sub(a,b) A::mul(a,k) A::fma(a,k,b)=a*k+b A::fms(a,k,b)=a*k-b A::neg(a),
    with k = A::cst(i, literal): entry i of the constant table dct64_kd (the double policy: a constant-bank operand of the
    DMUL / DFMA instead of two UMOVs per use) or the literal itself (the float policy: an immediate)."""
    n = p.n_in

    def cst(k):
        lit = _lit(k)
        idx = table.setdefault(lit, len(table))
        return f"A::cst({idx}, {lit})"

    lines = []
    lines.append(f"template <typename A>\n__device__ __forceinline__ void {name}(typename A::V (&x)[{n}]) {{")
    lines.append("  typedef typename A::V V;")

    def ref(i):
        return f"x[{i}]" if i < n else f"t{i}"

    for kind, dst, a, b, k in p.ops:
        if kind == "add":
            e = f"A::add({ref(a)}, {ref(b)})"
        elif kind == "sub":
            e = f"A::sub({ref(a)}, {ref(b)})"
        elif kind == "mul":
            e = f"A::mul({ref(a)}, {cst(k)})"
        elif kind == "fma":
            e = f"A::fma({ref(a)}, {cst(k)}, {ref(b)})"
        elif kind == "fms":
            e = f"A::fms({ref(a)}, {cst(k)}, {ref(b)})"
        lines.append(f"  const V t{dst} = {e};")
    for i, (vid, s) in enumerate(p.outputs):
        lines.append(f"  x[{i}] = {ref(vid) if s > 0 else 'A::neg(' + ref(vid) + ')'};")
    lines.append("}")
    return "\n".join(lines)


HEADER = """// GENERATED by tools/gen_dct64.py -- do not edit by hand.
// Straight-line 64-point orthonormal DCT-II (forward) and DCT-III (inverse), one block per
// thread, all 64 values register resident.  Replaces the FFTW-backed dct_fftw()/ifft_idct()
// of the reference (dct.c:55-103, dct.c:115-205; float twins in dct-float.c) for dn == 64.
// Operation counts: forward %(fwd)s ; inverse %(inv)s
#pragma once
"""


def main():
    out = os.path.join(os.path.dirname(__file__), "..", "dctz_b200", "csrc", "dct64_gen.cuh")
    if len(sys.argv) > 1:
        out = sys.argv[1]
    f = build_forward(64)
    i = build_inverse(64)
    txt = HEADER % {"fwd": op_counts(f), "inv": op_counts(i)}
    table = {}
    body = emit_function(f, "dct64_forward", table) + "\n\n" + emit_function(i, "dct64_inverse", table) + "\n"
    lits = sorted(table, key=table.get)
    txt += f"\n// the distinct constants of both flow graphs ({len(lits)}); see A::cst\n"
    txt += ("// (left uninitialised on the device and uploaded by the host once per context: with an initialiser in sight the compiler\n"
            "// folds dct64_kd[i] back into the literal)\n")
    txt += f"constexpr int DCT64_NK = {len(lits)};\n__constant__ double dct64_kd[DCT64_NK];\n"
    txt += "static const double dct64_kd_host[DCT64_NK] = {\n" + "".join(f"    {v},\n" for v in lits) + "};\n"
    txt += "\n" + body
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as fh:
        fh.write(txt)
    print("wrote", os.path.normpath(out), op_counts(f), op_counts(i))


if __name__ == "__main__":
    main()
