#!/usr/bin/env python3
"""Generate straight-line complex-FFT codelets with literal twiddles.

Writes aas_enhancement_b200/csrc/fft_codelets.cuh.  Each codelet is a force-inlined
function over two fixed-size float arrays (re, im) indexed only by compile-time constants,
so that nvcc keeps them in registers.  Trivial twiddles (1, -i, the eighth roots) are folded
at generation time; signs are tracked symbolically so no negations are emitted.

Forward transform convention: X[k] = sum_n x[n] exp(-2*pi*i*n*k/N).

    python tools/gen_fft_codelets.py            # regenerate the header
"""
from __future__ import annotations

import math
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "aas_enhancement_b200", "csrc", "fft_codelets.cuh")


class Emitter:
    def __init__(self):
        self.lines = []
        self.n = 0
        self.flops = 0

    def tmp(self, expr):
        name = f"v{self.n}"
        self.n += 1
        self.lines.append(f"    const float {name} = {expr};")
        return name


class R:
    """A real value: sign * variable-name (or exact zero when name is None)."""
    __slots__ = ("s", "v")

    def __init__(self, v, s=1):
        self.v, self.s = v, s

    def neg(self):
        return R(self.v, -self.s)


def radd(e, a, b):
    if a.v is None:
        return b
    if b.v is None:
        return a
    e.flops += 1
    if a.s > 0 and b.s > 0:
        return R(e.tmp(f"{a.v} + {b.v}"))
    if a.s > 0 and b.s < 0:
        return R(e.tmp(f"{a.v} - {b.v}"))
    if a.s < 0 and b.s > 0:
        return R(e.tmp(f"{b.v} - {a.v}"))
    return R(e.tmp(f"{a.v} + {b.v}"), -1)


def rsub(e, a, b):
    return radd(e, a, b.neg())


def lit(x):
    return f"{x:.9e}f"


def rmulc(e, a, c):
    """a * literal c"""
    if a.v is None or c == 0.0:
        return R(None)
    if c == 1.0:
        return a
    if c == -1.0:
        return a.neg()
    e.flops += 1
    s = a.s
    if c < 0:
        c, s = -c, -s
    return R(e.tmp(f"{a.v} * {lit(c)}"), s)


def rfma(e, a, c, b):
    """a * c + b with literal c (emitted so the compiler can contract to an FMA)."""
    if a.v is None or c == 0.0:
        return b
    if b.v is None:
        return rmulc(e, a, c)
    e.flops += 1
    # fold signs into the literal: result = sb * ( (sa*sb*c) * a + b )
    cc = c * a.s * b.s
    return R(e.tmp(f"fmaf({a.v}, {lit(cc)}, {b.v})"), b.s)


class C:
    __slots__ = ("re", "im")

    def __init__(self, re, im):
        self.re, self.im = re, im


def cadd(e, a, b):
    return C(radd(e, a.re, b.re), radd(e, a.im, b.im))


def csub(e, a, b):
    return C(rsub(e, a.re, b.re), rsub(e, a.im, b.im))


def cmul_mi(a):
    """multiply by -i: (re, im) -> (im, -re)"""
    return C(a.im, a.re.neg())


def ctwiddle(e, a, num, den):
    """a * exp(-2*pi*i*num/den) with folding of the trivial cases."""
    num %= den
    g = math.gcd(num, den)
    num, den = num // g, den // g
    if num == 0:
        return a
    if den == 2:            # -1
        return C(a.re.neg(), a.im.neg())
    if den == 4:
        return cmul_mi(a) if num == 1 else C(a.im.neg(), a.re)       # -i / +i
    ang = -2.0 * math.pi * num / den
    c, s = math.cos(ang), math.sin(ang)
    if den == 8:
        # |c| == |s| = r: (re*c - im*s, re*s + im*c)
        r = math.sqrt(0.5)
        sc = 1 if c > 0 else -1
        ss = 1 if s > 0 else -1
        # re' = r*(sc*re - ss*im) ; im' = r*(ss*re + sc*im)
        re_t = radd(e, R(a.re.v, a.re.s * sc), R(a.im.v, -a.im.s * ss))
        im_t = radd(e, R(a.re.v, a.re.s * ss), R(a.im.v, a.im.s * sc))
        return C(rmulc(e, re_t, r), rmulc(e, im_t, r))
    # general: 2 mul + 2 fma
    t1 = rmulc(e, a.im, -s)               # -im*s
    re_t = rfma(e, a.re, c, t1)           # re*c - im*s
    t2 = rmulc(e, a.im, c)                # im*c
    im_t = rfma(e, a.re, s, t2)           # re*s + im*c
    return C(re_t, im_t)


def dft2(e, x):
    return [cadd(e, x[0], x[1]), csub(e, x[0], x[1])]


def dft4(e, x):
    a = cadd(e, x[0], x[2])
    b = csub(e, x[0], x[2])
    c = cadd(e, x[1], x[3])
    d = cmul_mi(csub(e, x[1], x[3]))
    return [cadd(e, a, c), cadd(e, b, d), csub(e, a, c), csub(e, b, d)]


def fft(e, x):
    """Recursive mixed radix (4 where possible, else 2) decimation in time."""
    n = len(x)
    if n == 1:
        return x
    if n == 2:
        return dft2(e, x)
    if n == 4:
        return dft4(e, x)
    r = 4 if n % 4 == 0 and n > 8 else 2
    if n == 8:
        r = 2
    m = n // r
    subs = [fft(e, x[j::r]) for j in range(r)]          # r transforms of length m
    out = [None] * n
    for q in range(m):
        col = [ctwiddle(e, subs[j][q], j * q, n) for j in range(r)]
        y = dft4(e, col) if r == 4 else dft2(e, col)
        for s in range(r):
            out[q + m * s] = y[s]
    return out


def gen_codelet(n):
    e = Emitter()
    x = [C(R(f"xr[{i}]"), R(f"xi[{i}]")) for i in range(n)]
    # snapshot inputs so that in-place output is safe
    y = fft(e, x)
    body = []
    body.append(f"// complex {n}-point forward DFT, natural order in and out, in place")
    body.append(f"LMFB_HD void fft{n}(float (&xr)[{n}], float (&xi)[{n}]) {{")
    body.extend(e.lines)

    def ref(r):
        if r.v is None:
            return "0.0f"
        return r.v if r.s > 0 else f"-{r.v}"

    # outputs may alias inputs (names xr[i]); write through temporaries
    outs = []
    for k in range(n):
        outs.append((f"    const float or{k} = {ref(y[k].re)}; const float oi{k} = {ref(y[k].im)};"))
    body.extend(outs)
    for k in range(n):
        body.append(f"    xr[{k}] = or{k}; xi[{k}] = oi{k};")
    body.append("}")
    return "\n".join(body), e.flops


def main():
    parts = []
    parts.append("// GENERATED by tools/gen_fft_codelets.py -- do not edit by hand.")
    parts.append("#pragma once")
    parts.append("#ifndef LMFB_HD")
    parts.append("#  ifdef __CUDACC__")
    parts.append("#    define LMFB_HD __device__ __forceinline__")
    parts.append("#  else")
    parts.append("#    define LMFB_HD inline __attribute__((always_inline))")
    parts.append("#  endif")
    parts.append("#endif")
    parts.append("#include <math.h>")
    parts.append("namespace aas_lmfb {")
    for n in (32,):
        code, flops = gen_codelet(n)
        parts.append(f"// fft{n}: {flops} floating-point operations (fma counted once)")
        parts.append(code)
    # split twiddles for the real-FFT post-pass (theta = 2*pi*f/320) and the bin of every pass-2
    # step (k2) and output (k1): copied once per CTA into the shared-memory step table
    parts.append("#ifdef __CUDACC__\n#  define LMFB_CONST __constant__ const\n#else\n#  define LMFB_CONST static const\n#endif")
    rows_s, rows_c, rows_f = [], [], []
    for k2 in range(17):
        fs = [(96 * k1 + 65 * k2) % 160 for k1 in range(5)]
        rows_s.append(", ".join(lit(math.sin(2.0 * math.pi * f / 320.0)) for f in fs) + ", 0.0f, 0.0f, 0.0f")
        rows_c.append(", ".join(lit(math.cos(2.0 * math.pi * f / 320.0)) for f in fs) + ", 0.0f, 0.0f, 0.0f")
        rows_f.append(", ".join(str(f) for f in fs) + ", 0, 0, 0")
    parts.append("// [k2][k1] (rows padded to 8 words): split twiddles and bin f = (96*k1 + 65*k2) mod 160")
    body_s = " = {\n  {" + "},\n  {".join(rows_s) + "}};"
    body_c = " = {\n  {" + "},\n  {".join(rows_c) + "}};"
    body_f = " = {\n  {" + "},\n  {".join(rows_f) + "}};"
    parts.append("LMFB_CONST float kStepSin[17][8]" + body_s)
    parts.append("LMFB_CONST float kStepCos[17][8]" + body_c)
    parts.append("LMFB_CONST unsigned kStepBin[17][8]" + body_f)
    # host copies (the library builds the shared-memory table image on the host once per plan)
    parts.append("#ifdef __CUDACC__")
    parts.append("static const float kStepSinHost[17][8]" + body_s)
    parts.append("static const float kStepCosHost[17][8]" + body_c)
    parts.append("static const unsigned kStepBinHost[17][8]" + body_f)
    parts.append("#else\n#  define kStepSinHost kStepSin\n#  define kStepCosHost kStepCos\n#  define kStepBinHost kStepBin\n#endif")
    parts.append("}  // namespace aas_lmfb")
    with open(OUT, "w") as f:
        f.write("\n".join(parts) + "\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
