import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# Parity tolerance stated by BASELINE.json north_star / BASELINE.md 3:
#   |x - ref| <= 1e-5 * max(|ref|, 1)      (fp32; positions in metres, keypoints, gradients)
RTOL = 1e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_err(x, ref, row_scale=None, floor=1.0):
    """max |x - ref| / max(|ref|, floor).  `floor` (scalar or array like ref) is the magnitude below which the
    error is judged absolutely; 1 for quantities the north-star tolerance is stated on."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    if x.size == 0:
        return 0.0
    e = np.abs(x - ref) / np.maximum(np.abs(ref), floor)
    if row_scale is not None:
        e = e / np.asarray(row_scale, dtype=np.float64).reshape((-1,) + (1,) * (e.ndim - 1))
    return float(np.max(e))


def assert_parity(x, ref, what="", rtol=RTOL, row_scale=None, floor=1.0):
    e = rel_err(x, ref, row_scale, floor)
    assert np.isfinite(e) and e <= rtol, "%s: max |x-ref|/max(|ref|,1) = %.3e > %.1e" % (what, e, rtol)
    return e


def assert_parity_vs_reference(x, ref32, exact64, what="", rtol=RTOL, c=4.0):
    """Reference-relative bound (parity item): against the reference's own fp32 result `ref32`,
        |x - ref32| <= rtol * max(|ref32|, 1)  +  c * E_ref(pose),    E_ref(pose) = max_elements |ref32 - exact64|,
    where `exact64` is the float64 oracle on the same inputs.  For every normally placed pose the reference sits ~1e-7
    from exact arithmetic, so the second term is nothing and the plain north-star bound applies; for poses the generator
    drops onto the camera plane (clamp active, x/z ill-conditioned) the kernel is allowed c times the reference's OWN
    distance from the exact value of that pose -- not an a-priori error model.  Returns (max scaled error, number of
    poses that needed the second term)."""
    x = np.asarray(x, dtype=np.float64)
    ref32 = np.asarray(ref32, dtype=np.float64)
    exact64 = np.asarray(exact64, dtype=np.float64)
    assert x.shape == ref32.shape == exact64.shape, (x.shape, ref32.shape, exact64.shape)
    n = x.shape[0]
    e_ref = np.abs(ref32 - exact64).reshape(n, -1).max(axis=1).reshape((n,) + (1,) * (x.ndim - 1))
    plain = rtol * np.maximum(np.abs(ref32), 1.0)
    err = np.abs(x - ref32)
    ok = err <= plain + c * e_ref
    assert np.isfinite(err).all() and ok.all(), "%s: %d elements beyond rtol*max(|ref|,1) + %g*E_ref; worst %.3e (plain bound %.1e, E_ref %.3e)" % (
        what, int((~ok).sum()), float((err / plain).max()), rtol,
        float(e_ref.reshape(-1)[np.unravel_index(np.argmax(err / plain), err.shape)[0]]))
    needed = int((~(err <= plain)).reshape(n, -1).any(axis=1).sum())
    return float((err / (plain + c * e_ref)).max()), needed


def projection_conditioning(cam_xyz, world=None, cam_block=None, g_uv=None, kind="grad"):
    """Per-pose tolerance multiplier (>= 1) for quantities that pass through x/z.  It is exactly 1 for
    every pose in front of the camera at a normal distance, so those are held to the plain 1e-5 bound.

    First-order fp32 error model.  Camera-space coordinates are differences of world coordinates and
    the camera position, so they carry an absolute rounding error of about eps*|W| (eps = 2^-23,
    |W| = largest world / camera-translation coordinate).  Propagating it:
        uv_k          error ~ eps*|W| * f * 2 / |z_k|
        d/d(inputs)   error ~ 4 * eps*|W| * f * sum_k |g_uv_k|_1 / z_k^2  (the 16 per-joint terms ADD in
                      magnitude even when their sum cancels, so the bound is not relative to |sum|;
                      the factor 4 covers the radial/tangential distortion terms, whose derivative
                      reaches ~3x the pinhole one at the clamp edge |x/z| = 1)
    When a joint is within ~1 m of the camera plane (only the 10*tanh(randn) root mode produces such
    poses: the generator drops the skeleton onto the camera) this bound exceeds 1e-5 for ANY fp32
    implementation -- the reference's own torch result misses the float64 value by 1.2e-5 .. 1.7e-5 on
    stress poses 112802 / 32104 / 965556.  multiplier = max(1, bound / 1e-5)."""
    eps = 2.0 ** -23
    c = np.abs(np.asarray(cam_xyz, dtype=np.float64))
    z = np.maximum(c[..., 2], 1e-12)
    wmax = c.max(axis=(-1, -2))
    if world is not None:
        wmax = np.maximum(wmax, np.abs(np.asarray(world, dtype=np.float64)).max(axis=(-1, -2)))
    fmax = 2.3
    if cam_block is not None:
        cb = np.asarray(cam_block, dtype=np.float64).reshape(-1)
        fmax = float(np.abs(cb[7:9]).max())
        wmax = np.maximum(wmax, np.abs(cb[4:7]).max())
    if kind == "uv":
        bound = eps * wmax * fmax * 2.0 / z.min(axis=-1)
    else:
        gu = np.abs(np.asarray(g_uv, dtype=np.float64)).sum(axis=-1) if g_uv is not None else np.full(z.shape, 2.0)
        bound = 4.0 * eps * wmax * fmax * (gu / z ** 2).sum(axis=-1)
    return np.maximum(1.0, bound / RTOL)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    return load


@pytest.fixture(scope="session")
def c_oracle():
    import c_oracle as co
    co.build()
    return co
