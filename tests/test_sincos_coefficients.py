"""The forward kernel's accurate sincos (csrc/dhfk_device.cuh, DHFK_SINCOS_HALFTURN): its ten polynomial coefficients,
read out of the header, in an fp32 emulation of the exact instruction sequence (tools/sincos_reduction_study.py) against
float64 -- a guard against an edited digit, and the record of the error bound DESIGN.md quotes.  No GPU."""
import glob
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _header_coefficients():
    src = open(glob.glob(os.path.join(ROOT, "dh-aug*_b200", "csrc", "dhfk_device.cuh"))[0]).read()
    block = src[src.index("DHFK_SINCOS_HALFTURN) {"):]
    block = block[:block.index("#if DHFK_PACKED_V3")]
    get = lambda name: float(re.search(r"constexpr float %s = \(float\)\((-?[0-9.]+e[-+][0-9]+) \*" % name, block).group(1))
    return [get("S%d" % i) for i in range(5)], [get("C%d" % i) for i in range(5)]


def test_halfturn_sincos_error_bound():
    import sincos_reduction_study as study
    S, C = _header_coefficients()
    assert abs(S[0] - 1.0) < 1e-7 and abs(C[0] + 0.5) < 1e-7          # the leading terms of sin x / x and (cos x - 1) / x^2
    rng = np.random.RandomState(1)
    deg = np.concatenate([np.linspace(-720, 720, 400_001), rng.uniform(-180, 180, 400_000),
                          np.array([0.0, 90.0, -90.0, 180.0, -180.0, 270.0, 360.0, 1e5, -3e6])]).astype(np.float32)
    s, c = study.emulate_halfturn(deg, S, C)
    x = deg.astype(np.float64) * np.pi / 180.0
    assert np.abs(s - np.sin(x)).max() < 1.3e-7
    assert np.abs(c - np.cos(x)).max() < 1.3e-7
    # exact values where fp32 can hold them: the reduction is exact and sin(0) / cos(0) come out of the polynomials exactly
    s0, c0 = study.emulate_halfturn(np.array([0.0, 180.0, 360.0, -180.0], np.float32), S, C)
    assert np.all(s0 == 0.0) and np.all(np.abs(c0) == 1.0) and list(np.sign(c0)) == [1, -1, 1, -1]
    # the quarter-turn scheme it replaced stays the more accurate one; the new one is within 1 ulp of 1.0
    sa, ca = study.scheme_a(deg)
    assert np.abs(sa - np.sin(x)).max() < np.abs(s - np.sin(x)).max() < 2 ** -23 * 1.05
