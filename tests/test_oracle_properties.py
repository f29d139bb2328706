"""Domain properties of the float64 C oracle (beyond the reference goldens that pin it): translation equivariance,
360-degree periodicity, bone-length homogeneity, rigidity under the global rotation, and its analytic backward
against central finite differences of its own forward."""
import numpy as np
import pytest

from dhfk import synthetic, tables


@pytest.fixture(scope="module")
def case():
    inp = synthetic.gan_like(64, seed=5)
    return inp, tables.camera_block("S6", 1)


def test_translation_equivariance_and_periodicity(c_oracle, case):
    inp, blk = case
    o = c_oracle.forward(inp["ang"], inp["grot"], inp["bone"], inp["root"])["world16"]
    d = np.array([[0.25, -1.5, 2.0]], np.float32)
    o2 = c_oracle.forward(inp["ang"], inp["grot"], inp["bone"], inp["root"] + d)["world16"]
    assert np.abs(o2 - (o + d[:, None, :])).max() < 1e-6
    o3 = c_oracle.forward(inp["ang"] + 360.0, inp["grot"] - 360.0, inp["bone"], inp["root"])["world16"]
    assert np.abs(o3 - o).max() < 2e-5          # fp32 inputs: ang + 360 rounds at the 1e-5 degree level
    assert np.abs(o[:, 0] - inp["root"]).max() < 1e-6      # Hip is the chain origin


def test_bone_length_homogeneity_and_rigidity(c_oracle, case):
    inp, blk = case
    zero = np.zeros_like(inp["root"])
    o = c_oracle.forward(inp["ang"], inp["grot"], inp["bone"], zero)["world16"]
    o2 = c_oracle.forward(inp["ang"], inp["grot"], 2.0 * inp["bone"], zero)["world16"]
    assert np.abs(o2 - 2.0 * o).max() < 1e-6
    # the global rotation is rigid: pairwise joint distances do not depend on it
    g2 = np.random.RandomState(1).uniform(-180, 180, inp["grot"].shape).astype(np.float32)
    o3 = c_oracle.forward(inp["ang"], g2, inp["bone"], zero)["world16"]
    dist = lambda x: np.linalg.norm(x[:, :, None, :] - x[:, None, :, :], axis=-1)
    assert np.abs(dist(o) - dist(o3)).max() < 1e-6
    # bone lengths come out as given (used_16key_15bone_len_table order)
    pairs = np.array(tables.used_16key_15bone_len_table)
    bl = np.linalg.norm(o[:, pairs[:, 1]] - o[:, pairs[:, 0]], axis=-1)
    assert np.abs(bl - inp["bone"]).max() < 1e-6


@pytest.mark.parametrize("which", ["w", "wu", "wcu"])
def test_backward_against_finite_differences(c_oracle, case, which):
    inp, blk = case
    n = 6
    sub = {k: v[:n].astype(np.float32) for k, v in inp.items()}
    up = synthetic.upstream_grads(n, seed=9)
    gw = up["g_world"]; gc = up["g_cam"] if "c" in which else None; gu = up["g_uv"] if "u" in which else None

    def loss(ang, grot, root):
        o = c_oracle.forward(ang, grot, sub["bone"], root, blk)
        v = (o["world16"] * gw).sum(axis=(1, 2))
        if gc is not None:
            v = v + (o["cam"] * gc).sum(axis=(1, 2))
        if gu is not None:
            v = v + (o["uv"] * gu).sum(axis=(1, 2))
        return v

    b = c_oracle.backward(sub["ang"], sub["grot"], sub["bone"], sub["root"], blk, g_world=gw, g_cam=gc, g_uv=gu)
    # fp32 inputs: steps that are exactly representable around the sampled values
    for name, key, h, cols in (("g_ang", "ang", 2.0 ** -6, range(0, 33, 4)), ("g_grot", "grot", 2.0 ** -6, range(3)),
                               ("g_root", "root", 2.0 ** -10, range(3))):
        for c in cols:
            args = {k: sub[k].copy() for k in ("ang", "grot", "root")}
            hi = {k: v.copy() for k, v in args.items()}; lo = {k: v.copy() for k, v in args.items()}
            hi[key][:, c] += h; lo[key][:, c] -= h
            step = (hi[key][:, c].astype(np.float64) - lo[key][:, c].astype(np.float64))
            fd = (loss(hi["ang"], hi["grot"], hi["root"]) - loss(lo["ang"], lo["grot"], lo["root"])) / step
            ref = b[name][:, c]
            assert np.abs(fd - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max()), (name, c)
