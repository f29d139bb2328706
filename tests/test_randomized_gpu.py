"""Seeded randomized differential test: many small launches with random batch sizes (around the 32-row tile), random
output / gradient subsets, packed and strided inputs, all 28 H36M cameras, both angle distributions and all three trig
policies, each compared with the float64 C oracle at the north-star tolerance."""
import numpy as np
import pytest
import torch

from conftest import assert_parity, projection_conditioning

pytestmark = pytest.mark.gpu

SUBJECTS = ("S1", "S5", "S6", "S7", "S8", "S9", "S11")
TRIG = (dict(), dict(accurate_grad=True), dict(fast_trig=True))


def T(x, grad=False):
    return torch.tensor(np.asarray(x, dtype=np.float32), device="cuda:0", requires_grad=grad)


@pytest.mark.parametrize("seed", range(6))
def test_random_configurations_against_the_oracle(c_oracle, seed):
    import dhfk
    from dhfk import synthetic, tables
    rng = np.random.RandomState(1000 + seed)
    for trial in range(25):
        n = int(rng.choice([1, 2, 31, 32, 33, 63, 64, 65, 97, int(rng.randint(1, 3000))]))
        stress = bool(rng.randint(2))
        inp = synthetic.gan_like(n, seed=int(rng.randint(1 << 30)), angle_mode="stress" if stress else "gan")
        up = synthetic.upstream_grads(n, seed=int(rng.randint(1 << 30)))
        blk = tables.camera_block(SUBJECTS[rng.randint(7)], int(rng.randint(4)))
        trig = TRIG[rng.randint(3)]
        want_cam = bool(rng.randint(2))
        use = [bool(rng.randint(2)) for _ in range(3)]            # which of g_world / g_cam / g_uv carry a gradient
        if not any(use):
            use[0] = True
        use[1] = use[1] and want_cam
        if not any(use):
            use[2] = True
        strided = bool(rng.randint(2))
        if strided:                                               # the generator's layout: one [N,37] tensor, root from [N,35]
            g37 = torch.zeros(n, 37, device="cuda:0")
            g37[:, :33] = T(inp["ang"]); g37[:, 34:] = T(inp["grot"])
            g37.requires_grad_(True)
            ang, grot = g37[:, :33], g37[:, 34:]
            r35 = torch.zeros(n, 35, device="cuda:0")
            r35[:, 32:] = T(inp["root"])
            r35.requires_grad_(True)
            root = r35[:, 32:]
        else:
            ang, grot, root = T(inp["ang"], True), T(inp["grot"], True), T(inp["root"], True)
        world, cam, uv = dhfk.fk_project(ang, grot, T(inp["bone"]), root, blk, return_cam=want_cam, **trig)
        o = c_oracle.forward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk)
        tag = "seed %d trial %d n %d %s" % (seed, trial, n, trig)
        assert_parity(world.detach().cpu().numpy(), o["world16"], "world16 " + tag)
        ucond = projection_conditioning(o["cam"], o["world16"], blk, kind="uv")
        assert_parity(uv.detach().cpu().numpy(), o["uv"], "uv " + tag, row_scale=ucond)
        if want_cam:
            assert_parity(cam.detach().cpu().numpy(), o["cam"], "cam " + tag)
        loss = 0
        if use[0]:
            loss = loss + (world * T(up["g_world"])).sum()
        if use[1]:
            loss = loss + (cam * T(up["g_cam"])).sum()
        if use[2]:
            loss = loss + (uv * T(up["g_uv"])).sum()
        loss.backward()
        b = c_oracle.backward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk,
                              g_world=up["g_world"] if use[0] else None, g_cam=up["g_cam"] if use[1] else None,
                              g_uv=up["g_uv"] if use[2] else None, want_bone=False)
        cond = projection_conditioning(o["cam"], o["world16"], blk, up["g_uv"]) if use[2] else None
        if use[2]:        # poses with a point on the clamp edge within fp32 resolution: the mask is discontinuous there
            ratio = np.abs(o["cam"][..., :2] / o["cam"][..., 2:])
            band = 4e-6 * (1.0 + ratio) / np.abs(o["cam"][..., 2:])
            ok = ~(np.abs(ratio - 1) < band).any(axis=(1, 2))
        else:
            ok = np.ones(n, bool)
        if strided:
            g_ang, g_grot, g_root = g37.grad[:, :33], g37.grad[:, 34:], r35.grad[:, 32:]
            assert (g37.grad[:, 33] == 0).all() and (r35.grad[:, :32] == 0).all()
        else:
            g_ang, g_grot, g_root = ang.grad, grot.grad, root.grad
        for name, x, ref in (("g_ang", g_ang, b["g_ang"]), ("g_grot", g_grot, b["g_grot"]), ("g_root", g_root, b["g_root"])):
            assert_parity(x.cpu().numpy()[ok], ref[ok], name + " " + tag, row_scale=None if cond is None else cond[ok])
