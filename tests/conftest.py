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


def rel_err(x, ref, row_scale=None):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    if x.size == 0:
        return 0.0
    e = np.abs(x - ref) / np.maximum(np.abs(ref), 1.0)
    if row_scale is not None:
        e = e / np.asarray(row_scale, dtype=np.float64).reshape((-1,) + (1,) * (e.ndim - 1))
    return float(np.max(e))


def assert_parity(x, ref, what="", rtol=RTOL, row_scale=None):
    e = rel_err(x, ref, row_scale)
    assert np.isfinite(e) and e <= rtol, "%s: max |x-ref|/max(|ref|,1) = %.3e > %.1e" % (what, e, rtol)
    return e


def projection_conditioning(cam_xyz, kind="grad"):
    """Per-pose tolerance multiplier for quantities that pass through x/z (>= 1; exactly 1 for every
    pose that is in front of the camera at a normal distance).

    Camera-space coordinates carry an absolute fp32 rounding error of about eps*|X| (eps = 2^-23,
    |X| = largest coordinate).  First-order propagation:
        uv   = f(x/z):            error ~ eps * |X| / |z|          (per unclamped point)
        grad ~ g/z, g*x/z^2:      error ~ eps * |X| / z^2          (relative to an O(1) gradient)
    so a pose with a joint close to the camera plane (only the 10*tanh(randn) root mode produces
    these: the generator places the skeleton *inside* the camera) is ill-conditioned for ANY fp32
    implementation -- the reference's own torch result misses the float64 value by >1e-5 there
    (e.g. stress pose 32104: |X| = 10.9 m, z = -0.44 m, reference error 1.7e-5).  Such poses are held
    to 1e-5 * multiplier with
        grad: max(1, 0.5 * |X|max / zmin^2)        uv: max(1, 0.2 * |X|max / zmin)
    For the H36M set-up (|X| ~ 5 m, z ~ 5 m) both are 1 and the plain 1e-5 bound applies."""
    c = np.abs(np.asarray(cam_xyz, dtype=np.float64))
    zmin = np.maximum(c[..., 2].min(axis=-1), 1e-12)
    xmax = c.max(axis=(-1, -2))
    if kind == "uv":
        return np.maximum(1.0, 0.2 * xmax / zmin)
    return np.maximum(1.0, 0.5 * xmax / zmin ** 2)


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
