#!/usr/bin/env python
"""fp32 emulation of two accurate sincos(deg) schemes: (a) the kernels' current one -- reduce to |r| <= 45 deg, cephes
polynomials, swap + sign by quadrant -- and (b) reduce to |r| <= 90 deg, longer minimax-like polynomials, sign only.
Prints the max abs error of sin and cos against float64 over a dense sweep.  (Study tool; no GPU.)"""
import numpy as np

f32 = np.float32
D = np.pi / 180.0


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def scheme_a(deg):
    magic = f32(12582912.0)
    t = fma(deg, np.full_like(deg, f32(1.0 / 90.0)), np.full_like(deg, magic))
    n = t.view(np.int32).astype(np.int64)
    q = (t - magic).astype(f32)
    r = fma(q, np.full_like(deg, f32(-90.0)), deg)
    r2 = (r * r).astype(f32)
    S = [D, -1.6666654611e-1 * D ** 3, 8.3321608736e-3 * D ** 5, -1.9515295891e-4 * D ** 7]
    C = [-0.5 * D ** 2, 4.166664568298827e-2 * D ** 4, -1.388731625493765e-3 * D ** 6, 2.443315711809948e-5 * D ** 8]
    ps = fma(r2, np.full_like(deg, f32(S[3])), np.full_like(deg, f32(S[2])))
    ps = fma(r2, ps, np.full_like(deg, f32(S[1])))
    ps = fma(r2, ps, np.full_like(deg, f32(S[0])))
    sv = (r * ps).astype(f32)
    pc = fma(r2, np.full_like(deg, f32(C[3])), np.full_like(deg, f32(C[2])))
    pc = fma(r2, pc, np.full_like(deg, f32(C[1])))
    pc = fma(r2, pc, np.full_like(deg, f32(C[0])))
    cv = fma(r2, pc, np.ones_like(deg))
    odd = (n & 1) != 0
    so = np.where(odd, cv, sv)
    co = np.where(odd, sv, cv)
    s = np.where((n & 2) != 0, -so, so)
    c = np.where(((n + 1) & 2) != 0, -co, co)
    return s, c


def fit(kind, nterms, half_range_deg=90.0):
    """Least squares on Chebyshev nodes (close to minimax) of sin(x)/x resp. (cos(x)-1)/x^2 in u = x^2, x in radians."""
    k = np.arange(4000)
    x = np.cos((2 * k + 1) * np.pi / (2 * len(k))) * (half_range_deg * D)
    x = x[np.abs(x) > 1e-6]
    u = x * x
    y = np.sin(x) / x if kind == "sin" else (np.cos(x) - 1.0) / u
    A = np.vander(u, nterms, increasing=True)
    # weight for ABSOLUTE error of the final value: sin: error * x ; cos: error * u
    w = np.abs(x) if kind == "sin" else u
    coef, *_ = np.linalg.lstsq(A * w[:, None], y * w, rcond=None)
    return coef          # in radians^(2i)


def scheme_b(deg, ns, nc):
    return emulate_halfturn(deg, fit("sin", ns), fit("cos", nc))


def emulate_halfturn(deg, cs, cc):
    """The kernels' half-turn scheme in fp32 emulation with explicit coefficients (radian units, lowest order first)."""
    ns, nc = len(cs), len(cc)
    magic = f32(12582912.0)
    t = fma(deg, np.full_like(deg, f32(1.0 / 180.0)), np.full_like(deg, magic))
    n = t.view(np.int32).astype(np.int64)
    q = (t - magic).astype(f32)
    r = fma(q, np.full_like(deg, f32(-180.0)), deg)          # exact, |r| <= 90
    r2 = (r * r).astype(f32)
    S = [cs[i] * D ** (2 * i + 1) for i in range(ns)]        # degree units
    C = [cc[i] * D ** (2 * i + 2) for i in range(nc)]
    ps = np.full_like(deg, f32(S[-1]))
    for i in range(ns - 2, -1, -1):
        ps = fma(r2, ps, np.full_like(deg, f32(S[i])))
    sv = (r * ps).astype(f32)
    pc = np.full_like(deg, f32(C[-1]))
    for i in range(nc - 2, -1, -1):
        pc = fma(r2, pc, np.full_like(deg, f32(C[i])))
    cv = fma(r2, pc, np.ones_like(deg))
    odd = (n & 1) != 0
    return np.where(odd, -sv, sv), np.where(odd, -cv, cv)


def report(name, s, c, deg):
    x = deg.astype(np.float64) * D
    es, ec = np.abs(s - np.sin(x)), np.abs(c - np.cos(x))
    print("%-34s max |sin err| %.3e   max |cos err| %.3e   rms %.2e / %.2e" % (name, es.max(), ec.max(), np.sqrt((es ** 2).mean()),
                                                                         np.sqrt((ec ** 2).mean())))


if __name__ == "__main__":
    rng = np.random.RandomState(0)
    deg = np.concatenate([np.linspace(-720, 720, 2_000_001), rng.uniform(-180, 180, 2_000_000)]).astype(f32)
    report("(a) |r|<=45, 4+4 coefficients", *scheme_a(deg), deg)
    for ns, nc in ((5, 5), (6, 5), (6, 6), (5, 6)):
        report("(b) |r|<=90, %d+%d coefficients" % (ns, nc), *scheme_b(deg, ns, nc), deg)
