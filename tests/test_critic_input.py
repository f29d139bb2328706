"""SURVEY 8 f2 -- critic input transforms: left/right flip, root-centring and the KCS features
(models_Fk_GAN/Fk_discriminator.py:36-146, :269-377; special_operate.py:513-539; model_fk_gan_train.py:311-331).

CPU part: the C oracle (float64, complex-step derivatives) reproduces the goldens frozen from the unmodified
reference (outputs and autograd gradients).  GPU part: the fused kernels (forward / VJP / JVP) through the public
API against goldens and oracle, incl. WGAN-GP's double backward through the reference's own 3-D critic weights.
Tolerance |x - ref| <= 1e-5 * max(|ref|, 1) (north_star)."""
import numpy as np
import pytest
import torch

from conftest import assert_parity

COMBOS = [(0, 0), (1, 0), (0, 1), (1, 1)]


def _flags(c, f):
    return (1 if c else 0) | (2 if f else 0)


@pytest.mark.parametrize("centre,flip", COMBOS)
def test_c_oracle_matches_reference(golden, c_oracle, centre, flip):
    g = golden("critic")
    tag = "c%df%d" % (centre, flip)
    o = c_oracle.critic_forward(g["pose"], _flags(centre, flip), 30)
    assert_parity(o["pos"], g[tag + "_pos"], "pos")
    assert_parity(o["kcs"], g[tag + "_kcs"], "kcs")
    o15 = c_oracle.critic_forward(g["pose"], _flags(centre, flip), 15)
    assert_parity(o15["kcs"], g[tag + "_vkcs"], "video kcs")
    assert np.array_equal(o15["kcs"], o["kcs"][:, :15])
    b = c_oracle.critic_backward(g["pose"], g["g_pos"], g["g_kcs"], _flags(centre, flip))
    assert_parity(b, g[tag + "_g_pose"], "g_pose")
    assert_parity(c_oracle.critic_backward(g["pose"], None, g["g_kcs"], _flags(centre, flip)),
                  g[tag + "_g_pose_kcs_only"], "g_pose (kcs only)")
    assert_parity(c_oracle.critic_backward(g["pose"], None, g["g_kcs"][:, :15], _flags(centre, flip)),
                  g[tag + "_g_pose_vkcs_only"], "g_pose (video kcs only)")


def test_c_oracle_flip_and_adjoint_identity(golden, c_oracle):
    g = golden("critic")
    assert np.array_equal(c_oracle.flip(g["pose"]), g["flip3"].astype(np.float64))
    assert np.array_equal(c_oracle.flip(g["uv"]), g["flip2"].astype(np.float64))
    assert np.array_equal(c_oracle.flip(c_oracle.flip(g["uv"]).astype(np.float32)), g["uv"].astype(np.float64))
    # <J v, g> == <v, J^T g>
    v = np.random.RandomState(2).randn(*g["pose"].shape).astype(np.float32)
    for fl in (0, 3):
        t = c_oracle.critic_jvp(g["pose"], v, fl, 30)
        b = c_oracle.critic_backward(g["pose"], g["g_pos"], g["g_kcs"], fl)
        lhs = (t["pos"] * g["g_pos"]).sum() + (t["kcs"] * g["g_kcs"]).sum()
        rhs = (b * v).sum()
        assert abs(lhs - rhs) <= 1e-9 * max(abs(lhs), 1.0)


def test_kcs_invariances_on_oracle(golden, c_oracle):
    """Domain properties: features do not move under translation; the flip permutes them left<->right."""
    g = golden("critic")
    k0 = c_oracle.critic_forward(g["pose"], 0, 30)["kcs"]
    k1 = c_oracle.critic_forward(g["pose"], 1, 30)["kcs"]
    assert np.abs(k0 - k1).max() < 1e-6
    k2 = c_oracle.critic_forward(g["pose"], 2, 30)["kcs"]
    pair_swap = [1, 0, 3, 2, 4, 6, 5, 7, 8, 10, 9, 12, 11, 14, 13]
    bone_swap = [1, 0, 3, 2, 5, 4, 6, 7, 9, 8, 11, 10, 13, 12, 14]
    assert np.abs(k2[:, :15] - k0[:, pair_swap]).max() < 1e-12
    assert np.abs(k2[:, 15:] - k0[:, 15:][:, bone_swap]).max() < 1e-12


# ---------------------------------------------------------------------------------------- GPU
def T(x, grad=False):
    return torch.tensor(np.asarray(x, dtype=np.float32), device="cuda:0", requires_grad=grad)


@pytest.mark.gpu
@pytest.mark.parametrize("centre,flip", COMBOS)
def test_kernels_match_reference_golden(golden, centre, flip):
    import dhfk
    g = golden("critic")
    tag = "c%df%d" % (centre, flip)
    x = T(g["pose"], True)
    pos, kcs = dhfk.critic_input(x, centre=bool(centre), flip=bool(flip), kcs_cols=30)
    assert_parity(pos.detach().cpu().numpy(), g[tag + "_pos"], "pos")
    assert_parity(kcs.detach().cpu().numpy(), g[tag + "_kcs"], "kcs")
    ((pos * T(g["g_pos"])).sum() + (kcs * T(g["g_kcs"])).sum()).backward()
    assert_parity(x.grad.cpu().numpy(), g[tag + "_g_pose"], "g_pose")
    x2 = T(g["pose"], True)
    k2 = dhfk.critic_input(x2, centre=bool(centre), flip=bool(flip), kcs_cols=30, return_pos=False)
    assert torch.equal(k2, kcs)
    (k2 * T(g["g_kcs"])).sum().backward()
    assert_parity(x2.grad.cpu().numpy(), g[tag + "_g_pose_kcs_only"], "g_pose (kcs only)")
    x3 = T(g["pose"], True)
    p3, k3 = dhfk.critic_input(x3, centre=bool(centre), flip=bool(flip), kcs_cols=15)
    assert_parity(k3.detach().cpu().numpy(), g[tag + "_vkcs"], "video kcs")
    (k3 * T(g["g_kcs"][:, :15])).sum().backward()
    assert_parity(x3.grad.cpu().numpy(), g[tag + "_g_pose_vkcs_only"], "g_pose (video kcs only)")
    # positions only (kcs_cols=0): gradient is the transpose of centre/flip
    x4 = T(g["pose"], True)
    p4 = dhfk.critic_input(x4, centre=bool(centre), flip=bool(flip), kcs_cols=0)
    assert torch.equal(p4, pos)


@pytest.mark.gpu
def test_reference_shaped_functions_and_flip(golden):
    import dhfk
    from dhfk import Fk_discriminator as fd
    g = golden("critic")
    dev = torch.device("cuda:0")
    k = fd.special_KCS_Input_transform(T(g["pose"]).view(-1, 48), dev)          # the critic passes [N,48]
    assert k.shape == (133, 30)
    assert_parity(k.cpu().numpy(), g["c0f0_kcs"], "special_KCS_Input_transform")
    vk = fd.video_mode_special_KCS_Input_transform(T(g["pose"]), dev)
    assert_parity(vk.cpu().numpy(), g["c0f0_vkcs"], "video_mode_special_KCS_Input_transform")
    assert np.array_equal(dhfk.flip_pose(T(g["pose"])).cpu().numpy(), g["flip3"])   # bit-exact: a permutation + sign
    assert np.array_equal(dhfk.flip_pose(T(g["uv"])).cpu().numpy(), g["flip2"])
    u = T(g["uv"], True)
    (dhfk.flip_pose(u) * T(g["flip2"])).sum().backward()
    assert np.array_equal(u.grad.cpu().numpy(), g["uv"])                              # flip is its own transpose
    c, f = fd.critic_views(T(g["pose"]))
    assert_parity(c.cpu().numpy(), g["c1f0_pos"], "centred")
    assert_parity(f.cpu().numpy(), g["c1f1_pos"], "centred+flipped")


class _ResNet(torch.nn.Module):      # special_operate.py:490-510
    def __init__(self, d):
        super().__init__()
        self.fc1, self.fc2, self.relu = torch.nn.Linear(d, d), torch.nn.Linear(d, d), torch.nn.ReLU(True)

    def forward(self, x):
        out = self.fc2(self.relu(self.fc1(x)))
        out += x
        return self.relu(out)


class _Critic3D(torch.nn.Module):
    """Same layers / parameter names as the reference Fk_3D_Discriminator (Fk_discriminator.py:149-206); the KCS
    branch input comes from the native kernel."""

    def __init__(self, d):
        super().__init__()
        nn = torch.nn
        self.previous = nn.Sequential(nn.Linear(48, d), nn.ReLU(True))
        self.block1, self.block2, self.block3 = _ResNet(d), _ResNet(d), _ResNet(d)
        self.special_KCS_previous = nn.Sequential(nn.Linear(30, d), nn.ReLU(True))
        self.special_KCS_block1, self.special_KCS_block2, self.special_KCS_block3 = _ResNet(d), _ResNet(d), _ResNet(d)
        self.merge_previous = nn.Sequential(nn.Linear(2 * d, 100), nn.ReLU(True))
        self.merge_block1 = _ResNet(100)
        self.output = nn.Linear(100, 1)

    def forward(self, inp):
        from dhfk import Fk_discriminator as fd
        k = fd.special_KCS_Input_transform(torch.clone(inp), inp.device).contiguous().view(-1, 30)
        k = self.special_KCS_block3(self.special_KCS_block2(self.special_KCS_block1(self.special_KCS_previous(k))))
        p = inp.contiguous().view(-1, 48)
        p = self.block3(self.block2(self.block1(self.previous(p))))
        out = self.merge_block1(self.merge_previous(torch.cat((k, p), dim=-1)))
        return self.output(out)


@pytest.mark.gpu
def test_wgan_gp_double_backward_matches_reference_critic(golden):
    """calc_gradient_penalty (Fk_discriminator.py:208-233) on the reference critic's seeded weights: the penalty
    and every parameter gradient, which flow through the KCS transform's backward with create_graph=True."""
    g = golden("critic")
    D = _Critic3D(32).cuda()
    D.load_state_dict({k[len("d3d_w_"):]: T(v) for k, v in g.items() if k.startswith("d3d_w_")})
    real, fake = T(g["d3d_real"]), T(g["d3d_fake"])
    assert_parity(D(real).detach().cpu().numpy(), g["d3d_out_real"], "D(real)")
    B = real.shape[0]
    alpha = T(g["d3d_alpha"]).expand(B, 48)
    inter = (alpha * real.view(B, -1) + (1 - alpha) * fake.view(B, -1)).detach().requires_grad_(True)
    out = D(inter)
    (grads,) = torch.autograd.grad(out, inter, grad_outputs=torch.ones_like(out), create_graph=True,
                                   retain_graph=True, only_inputs=True)
    gp = ((grads.norm(2, dim=1) - 1) ** 2).mean() * 10
    D.zero_grad()
    gp.backward()
    assert abs(gp.item() - float(g["d3d_gp"][0])) <= 1e-5 * max(1.0, abs(float(g["d3d_gp"][0])))
    checked = 0
    for name, prm in D.named_parameters():
        ref = g["d3d_g_" + name]
        got = prm.grad.cpu().numpy() if prm.grad is not None else np.zeros_like(ref)
        scale = max(1.0, float(np.abs(ref).max()))
        assert np.abs(got - ref).max() <= 2e-5 * scale, name     # sums over 40 poses x 32-100 units in fp32
        checked += 1
    assert checked == len(g["d3d_param_names"])
    assert np.abs(g["d3d_g_special_KCS_previous.0.weight"]).max() > 0.1          # the KCS branch really contributes


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 31, 33, 4096 + 7, 262144])
def test_kernels_match_oracle_ragged_and_large(c_oracle, n):
    import dhfk
    from dhfk import synthetic
    d = synthetic.gan_like(max(n, 1), seed=11)
    pose = c_oracle.forward(d["ang"], d["grot"], d["bone"], d["root"])["world16"].astype(np.float32)[:n]
    rng = np.random.RandomState(n % 89)
    gp, gk, v = (rng.randn(n, 16, 3).astype(np.float32), rng.randn(n, 30).astype(np.float32),
                 rng.randn(n, 16, 3).astype(np.float32))
    for fl in (0, 3):
        x = T(pose, True)
        pos, kcs = dhfk.critic_input(x, centre=bool(fl & 1), flip=bool(fl & 2), kcs_cols=30)
        assert pos.shape == (n, 16, 3) and kcs.shape == (n, 30)
        if n == 0:
            continue
        ref = c_oracle.critic_forward(pose, fl, 30)
        assert_parity(pos.detach().cpu().numpy(), ref["pos"], "pos")
        assert_parity(kcs.detach().cpu().numpy(), ref["kcs"], "kcs")
        gpt, gkt = T(gp, True), T(gk, True)
        (gx,) = torch.autograd.grad((pos * gpt).sum() + (kcs * gkt).sum(), x, create_graph=True)
        # J^T g: entries are sums of up to ~20 terms of size |g|/|bone| ~ 10 -> judge relative to the row scale
        refb = c_oracle.critic_backward(pose, gp, gk, fl)
        scale = np.maximum(1.0, np.abs(refb).max(axis=(1, 2)))
        assert_parity(gx.detach().cpu().numpy(), refb, "g_pose", row_scale=scale)
        # derivative of the VJP w.r.t. the upstream gradients = JVP kernel
        t_pos, t_kcs = torch.autograd.grad((gx * T(v)).sum(), (gpt, gkt))
        reft = c_oracle.critic_jvp(pose, v, fl, 30)
        assert_parity(t_pos.cpu().numpy(), reft["pos"], "t_pos")
        tscale = np.maximum(1.0, np.abs(reft["kcs"]).max(axis=1))
        assert_parity(t_kcs.cpu().numpy(), reft["kcs"], "t_kcs", row_scale=tscale)


@pytest.mark.gpu
def test_input_shapes_devices_and_leading_dims(golden):
    """[N,48] rows (what the critic's interpolates look like), CPU tensors (moved to the device like every other entry
    point) and clip-shaped [B,F,16,d] inputs of the flip."""
    import dhfk
    g = golden("critic")
    x48 = T(g["pose"]).view(-1, 48).clone().requires_grad_(True)
    pos, kcs = dhfk.critic_input(x48, centre=True, kcs_cols=30)
    assert pos.shape == (133, 16, 3)
    assert_parity(kcs.detach().cpu().numpy(), g["c1f0_kcs"], "kcs from [N,48]")
    (kcs * T(g["g_kcs"])).sum().backward()
    assert x48.grad.shape == (133, 48)
    assert_parity(x48.grad.view(-1, 16, 3).cpu().numpy(), g["c1f0_g_pose_kcs_only"], "g_pose [N,48]")
    xc = torch.tensor(g["pose"], requires_grad=True)                       # CPU leaf
    k = dhfk.critic_input(xc, kcs_cols=15, return_pos=False)
    assert k.is_cuda
    (k * T(g["g_kcs"][:, :15])).sum().backward()
    assert not xc.grad.is_cuda
    assert_parity(xc.grad.numpy(), g["c0f0_g_pose_vkcs_only"], "g_pose on the CPU leaf")
    clip2 = T(g["uv"][:132]).view(12, 11, 16, 2)
    assert np.array_equal(dhfk.flip_pose(clip2).cpu().numpy().reshape(-1, 16, 2), g["flip2"][:132])
    clip3 = T(g["pose"][:132]).view(4, 33, 16, 3)
    assert np.array_equal(dhfk.flip_pose(clip3).cpu().numpy().reshape(-1, 16, 3), g["flip3"][:132])
    with pytest.raises(ValueError):
        dhfk.critic_input(T(g["pose"]), kcs_cols=7)
