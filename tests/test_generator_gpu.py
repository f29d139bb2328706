"""generator epilogue folded into the kernel prologue (dhfk_generator_forward/backward) and the
reference-shaped Fk_Generator / Video_Fk_Generator built on it, against goldens produced by the reference's own
generator code and against the float64 oracle."""
import argparse

import numpy as np
import pytest
import torch

from conftest import assert_parity

pytestmark = pytest.mark.gpu


def T(x, grad=False):
    return torch.tensor(np.asarray(x, dtype=np.float32), device="cuda", requires_grad=grad)


def _scaled_bone(bone, scaler):
    from dhfk import tables
    grp = tables.BONE_SCALER_GROUP
    return (bone * np.where(grp[None, :] >= 0, 1.0 + scaler[:, np.maximum(grp, 0)], 1.0)).astype(np.float32)


@pytest.mark.parametrize("trig", [dict(accurate_grad=True), {}, dict(fast_trig=True)], ids=["accurate", "default", "mufu"])
@pytest.mark.parametrize("tag,pre", [("single", True), ("single_nopre", False), ("video", True)])
def test_fused_generator_epilogue_matches_reference(golden, tag, pre, trig):
    import dhfk
    g = golden("generator")
    raw = T(g[tag + "_raw"].reshape(-1, 35), True)
    bone = T(_scaled_bone(g[tag + "_bone"], g[tag + "_scaler"]))
    world = dhfk.generator_fk(raw, bone, use_pre_angle=pre, **trig)
    assert_parity(world.detach().cpu().numpy(), g[tag + "_fake"].reshape(-1, 16, 3), "fake")
    (world * T(g[tag + "_g_fake"].reshape(-1, 16, 3))).sum().backward()
    assert_parity(raw.grad.cpu().numpy(), g[tag + "_d_raw"].reshape(-1, 35), "d_raw")
    assert torch.all(raw.grad[:, 31] == 0)


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 4608])
def test_generator_mode_with_camera_vs_oracle(c_oracle, n):
    """Generator mode + camera + projection outputs and gradients (what a fused GAN step uses), ragged sizes,
    and a strided raw-output view (the video generator's [B, F*35] tensor viewed as [B*F, 35] is packed; a
    [N, 40] buffer exercises the gather path)."""
    import dhfk
    from dhfk import synthetic, tables
    rng = np.random.RandomState(n)
    raw_np = (rng.randn(n, 35) * 0.8).astype(np.float32)
    raw_np[:, 32:35] = rng.uniform(-0.1, 0.1, (n, 3)); raw_np[:, 34] += 0.1     # keep roots near (0,0,1) m
    bone = synthetic.gan_like(n, seed=n + 1)["bone"]
    up = synthetic.upstream_grads(n, seed=n + 2)
    blk = tables.camera_block("S6", 1)
    half, mid = tables.generator_slot_scale(True)
    wide = torch.zeros(n, 40, device="cuda"); wide[:, :35] = T(raw_np); wide.requires_grad_(True)
    world, cam, uv = dhfk.generator_fk(wide[:, :35], T(bone), cam=blk, return_cam=True, return_uv=True)
    ((world * T(up["g_world"])).sum() + (cam * T(up["g_cam"])).sum() + (uv * T(up["g_uv"])).sum()).backward()
    o = c_oracle.gen_forward(raw_np, bone, half, mid, blk)
    d = c_oracle.gen_backward(raw_np, bone, half, mid, blk, g_world=up["g_world"], g_cam=up["g_cam"], g_uv=up["g_uv"])
    assert_parity(world.detach().cpu().numpy(), o["world16"], "world"); assert_parity(cam.detach().cpu().numpy(), o["cam"], "cam")
    assert_parity(uv.detach().cpu().numpy(), o["uv"], "uv")
    gw = wide.grad.cpu().numpy()
    # The 1e-5 bar is stated on the gradients of the FK inputs (angles, root).  d/d(raw) = that gradient times the
    # exact chain factor half*sech^2 (up to 180) or 10*sech^2, so the absolute floor scales with the factor.
    floor = np.maximum(1.0, c_oracle.gen_chain_factor(raw_np, half, mid))
    assert_parity(gw[:, :35], d, "d_raw", floor=floor)
    assert np.all(gw[:, 35:] == 0)


def _args(B, F, scaler_mode="different", pre=True):
    return argparse.Namespace(batch_size=B, random_seed=0, single_or_multi_train_mode="multi" if F > 1 else "single",
                              architecture="3,3", GAN_OUTPUT_DIM=35, Gen_DenseDim=16, GAN_whether_use_preAngle=pre,
                              whether_use_RT=True, bone_len_scaler=scaler_mode, record_all_picture=False, checkpoint="/tmp")


class _Feed(torch.nn.Module):
    def __init__(self, raw):
        super().__init__()
        self.raw = raw

    def forward(self, x):
        return self.raw * 1.0


def test_reference_shaped_generators(golden):
    """dhfk.Fk_generator.{Fk_Generator, Video_Fk_Generator}: same constructor / attributes / RNG use as the
    reference classes; fed with the golden raw outputs they reproduce the reference's fake poses and gradients."""
    from dhfk import Forward_Kinematics_DH_Model
    from dhfk.Fk_generator import Fk_Generator, Video_Fk_Generator
    g = golden("generator")
    dev = torch.device("cuda")
    # single frame: scaler drawn by torch.randint on the global CPU generator (Fk_generator.py:197)
    B = 70
    args = _args(B, 1)
    G = Fk_Generator(Forward_Kinematics_DH_Model(args, ["S1"], None), args, dev).to(dev)
    assert set(G.state_dict()) >= {"preprocess.0.weight", "block1.fc1.weight", "block3.fc2.bias", "deconv_out.weight"}
    raw = T(g["single_raw"], True)
    G.deconv_out = _Feed(raw)
    G.boneLength = T(g["single_bone"])
    torch.manual_seed(77)
    fake = G(torch.zeros(B, 128, device=dev))
    assert fake.shape == (B, 48)
    assert_parity(fake.detach().cpu().numpy(), g["single_fake"], "fake")
    (fake * T(g["single_g_fake"])).sum().backward()
    assert_parity(raw.grad.cpu().numpy(), g["single_d_raw"], "d_raw")
    # video: scaler from FK_DH_Class.random, repeated over the frames of a clip (Fk_generator.py:383-390)
    B, F = 4, 9
    args = _args(B, F)
    fk = Forward_Kinematics_DH_Model(args, ["S1"], None)
    G = Video_Fk_Generator(F, fk, args, dev).to(dev)
    assert G.deconv_out.out_features == F * 35
    raw = T(g["video_raw"], True)
    G.deconv_out = _Feed(raw)
    G.boneLength = T(g["video_bone"])
    fk.random = np.random.RandomState(5)
    fake = G(torch.zeros(B, 128, device=dev))
    assert fake.shape == (B, F, 48)
    assert_parity(fake.detach().cpu().numpy(), g["video_fake"], "video fake")
    (fake * T(g["video_g_fake"])).sum().backward()
    assert_parity(raw.grad.cpu().numpy(), g["video_d_raw"], "video d_raw")
    # bone lengths from real poses (GAN_generator_get_bone_length, :107-111)
    from dhfk.Fk_generator import bone_vectors_to_lengths
    g133 = golden("gan133")
    L = bone_vectors_to_lengths(T(g133["world16"]))
    assert_parity(L.cpu().numpy(), g133["bone"], "bone lengths from poses", rtol=2e-6)


def test_generator_mode_full_size_1m(c_oracle):
    import dhfk
    from dhfk import synthetic, tables
    n = 1 << 20
    rng = np.random.RandomState(3)
    raw_np = (rng.randn(n, 35) * 0.9).astype(np.float32)
    raw_np[:, 32:35] = rng.uniform(-0.1, 0.1, (n, 3)); raw_np[:, 34] += 0.1
    bone = synthetic.gan_like(n, seed=4)["bone"]
    up = synthetic.upstream_grads(n, seed=5)
    blk = tables.camera_block("S1", 0)
    half, mid = tables.generator_slot_scale(True)
    raw = T(raw_np, True)
    world, uv = dhfk.generator_fk(raw, T(bone), cam=blk, return_uv=True)
    ((world * T(up["g_world"])).sum() + (uv * T(up["g_uv"])).sum()).backward()
    o = c_oracle.gen_forward(raw_np, bone, half, mid, blk)
    d = c_oracle.gen_backward(raw_np, bone, half, mid, blk, g_world=up["g_world"], g_uv=up["g_uv"])
    e = [assert_parity(world.detach().cpu().numpy(), o["world16"], "world"),
         assert_parity(uv.detach().cpu().numpy(), o["uv"], "uv"),
         assert_parity(raw.grad.cpu().numpy(), d, "d_raw", floor=np.maximum(1.0, c_oracle.gen_chain_factor(raw_np, half, mid)))]
    print("\n[1M generator mode] max rel err world/uv/d_raw = %s" % " ".join("%.2e" % x for x in e))
