"""BASELINE configs 3 / 4 as parity cases: one generator step of the DH-AUG GAN (dense 256 MLPs) with the
native FK + camera kernels behind the reference-shaped interface, against the same step computed on the CPU
with the torch port of the reference (oracle/torch_port.py) and identical weights / noise.

Generator glue mirrors Fk_generator.py:114-259 (tanh, x10 root, 31 -> 37 slot scatter, per-slot range map,
bone length x (1 + scaler)); the critics are stand-in dense-256 MLPs (the real ones stay on cuBLAS torch and
are out of scope) -- what is checked is that the gradient reaching every generator parameter through
FK -> gather -> world->camera -> projection is the same."""
import argparse

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

ZERO_SLOTS = (4, 9, 22, 23, 28, 33)
BONES = ("left_small_leg_len", "right_small_leg_len", "left_big_leg_len", "right_big_leg_len", "left_hip_len",
         "right_hip_len", "waist_len", "thorax_len", "left_shoulder_len", "right_shoulder_len", "left_big_arm_len",
         "right_big_arm_len", "left_small_arm_len", "right_small_arm_len", "neck_len")
IDX16 = [0, 1, 2, 3, 6, 7, 8, 12, 13, 15, 17, 18, 19, 25, 26, 27]


class Res(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.a, self.b = nn.Linear(d, d), nn.Linear(d, d)

    def forward(self, x):
        return torch.relu(x + self.b(torch.relu(self.a(x))))


def mlp(i, o, d=256):
    return nn.Sequential(nn.Linear(i, d), nn.ReLU(), Res(d), Res(d), Res(d), nn.Linear(d, o))


def generator_glue(out35, frames):
    """Fk_generator.py:121-168 / :310-357: returns generator_angle [N,37], root [N,3] (or [B,F,3])."""
    from dhfk import tables
    out = out35.view(-1, 35)
    ang31 = torch.tanh(out[:, :-3])
    root = torch.tanh(out[:, -3:]) * 10.0
    cols, k = [], 0
    for i in range(37):
        if i in ZERO_SLOTS:
            cols.append(torch.zeros_like(ang31[:, 0]))
        else:
            cols.append(ang31[:, k]); k += 1
    g = torch.stack(cols, 1)
    rng = np.concatenate([tables.GAN_ANGLE_RANGE, tables.GAN_GLOBAL_ROT_RANGE]).astype(np.float32)
    lo = torch.as_tensor(rng[:, 0], device=g.device); hi = torch.as_tensor(rng[:, 1], device=g.device)
    g = g * ((hi - lo) / 2) + (hi + lo) / 2
    if frames > 1:
        root = root.view(-1, frames, 3)
    return g, root


def fk_kwargs(g, bone, root):
    kw = dict(right_leg_joints_angle=g[:, 0:5], left_leg_joints_angle=g[:, 5:10], body_joints_angle=g[:, 10:23],
              right_hand_joints_angle=g[:, 23:28], left_hand_joints_angle=g[:, 28:33],
              generator_global_rot_3d_pos_angle=g[:, -3:], root_3d_pos=root)
    for i, name in enumerate(BONES):
        kw[name] = bone[:, i]
    return kw


def gan_generator_step(device, G, D3, D2, noise, bone, blk, frames, native):
    G, D3, D2 = G.to(device), D3.to(device), D2.to(device)
    for m in (G, D3, D2):
        m.zero_grad()
    out = G(noise.to(device))
    g, root = generator_glue(out, frames)
    bone = bone.to(device)
    n = g.shape[0]
    blk_t = torch.as_tensor(blk, device=device)
    if native:
        from dhfk import Forward_Kinematics_DH_Model, camera
        args = argparse.Namespace(batch_size=n // frames, random_seed=0,
                                  single_or_multi_train_mode="multi" if frames > 1 else "single", architecture="3,3")
        w32 = Forward_Kinematics_DH_Model(args, ["S1"], None).change_3d_joint_angle(**fk_kwargs(g, bone, root))
        fake = w32[:, IDX16].view(-1, 16, 3)
        cam = camera.GAN_torch_world_to_camera(fake, R=blk_t[0:4].view(1, 4), t=blk_t[4:7].view(1, 3))
        uv = camera.project_to_2d(cam, blk_t[7:16].view(1, 9).repeat(n, 1))
    else:
        import torch_port
        w32 = torch_port.RefFKPort(n).fk32(g[:, :33], g[:, -3:], bone, root.reshape(-1, 3))
        fake = w32[:, IDX16].view(-1, 16, 3)
        cam = torch_port.world_to_camera(fake, blk_t[0:4].view(1, 4), blk_t[4:7].view(1, 3))
        uv = torch_port.project_to_2d(cam, blk_t[7:16].view(1, 9).repeat(n, 1))
    centred = fake - fake[:, :1]                                       # model_fk_gan_train.py:440
    loss = D3(centred.reshape(n, 48)).mean() * 1.0 + D2(uv.reshape(n, 32)).mean() * 0.2   # :474
    loss.backward()
    grads = [p.grad.detach().cpu().double() for p in G.parameters()]
    return loss.item(), grads, fake.detach().cpu(), uv.detach().cpu()


@pytest.mark.parametrize("batch,frames", [(1024, 1), (512, 9)], ids=["cfg3_single_b1024", "cfg4_video_b512x9"])
def test_generator_step_gradients_match_reference_path(batch, frames):
    from dhfk import synthetic, tables
    torch.manual_seed(0)
    G, D3, D2 = mlp(128, 35 * frames), mlp(48, 1), mlp(32, 1)
    # keep the roots in the camera's field of view so the projection gradients are well conditioned
    with torch.no_grad():
        last = G[-1]
        last.weight.mul_(0.05)
        b = last.bias.view(frames, 35)
        b[:, -3:] = torch.tensor([0.0, 0.0, 0.1])                   # 10*tanh(0.1) ~ 1 m above the floor
    noise = torch.randn(batch, 128)
    bone = torch.tensor(synthetic.gan_like(batch * frames, seed=4)["bone"])
    blk = tables.camera_block("S1", 0)
    torch.set_num_threads(8)
    l_ref, g_ref, fake_ref, uv_ref = gan_generator_step(torch.device("cpu"), G, D3, D2, noise, bone, blk, frames, False)
    l_nat, g_nat, fake_nat, uv_nat = gan_generator_step(torch.device("cuda"), G, D3, D2, noise, bone, blk, frames, True)
    assert abs(l_ref - l_nat) <= 1e-5 * max(1.0, abs(l_ref))
    assert (fake_ref - fake_nat).abs().max() <= 1e-5 * max(1.0, fake_ref.abs().max().item())
    assert (uv_ref - uv_nat).abs().max() <= 1e-5
    scale = max(g.abs().max().item() for g in g_ref)
    assert scale > 0
    for a, b in zip(g_ref, g_nat):
        # cuBLAS vs CPU GEMM rounding differs too; 1e-4 of the largest generator gradient
        assert (a - b).abs().max().item() <= 1e-4 * scale
