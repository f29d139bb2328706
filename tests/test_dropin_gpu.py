"""The reference-shaped surface (class + camera functions) on the GPU: same call signatures as
run_Fk_GAN.py / Fk_generator.py / model_fk_gan_train.py use, checked against the reference's goldens."""
import argparse

import numpy as np
import pytest
import torch

from conftest import assert_parity, projection_conditioning

pytestmark = pytest.mark.gpu

BONES = ("left_small_leg_len", "right_small_leg_len", "left_big_leg_len", "right_big_leg_len", "left_hip_len",
         "right_hip_len", "waist_len", "thorax_len", "left_shoulder_len", "right_shoulder_len", "left_big_arm_len",
         "right_big_arm_len", "left_small_arm_len", "right_small_arm_len", "neck_len")


def T(x, grad=False):
    return torch.tensor(np.asarray(x, dtype=np.float32), device="cuda", requires_grad=grad)


def _model(batch=16, mode="single", arch="3,3,3"):
    from dhfk import Forward_Kinematics_DH_Model
    args = argparse.Namespace(batch_size=batch, random_seed=0, single_or_multi_train_mode=mode, architecture=arch)
    return Forward_Kinematics_DH_Model(args, ["S1"], None)


def _kwargs(gen, bone, root):
    kw = dict(right_leg_joints_angle=gen[:, 0:5], left_leg_joints_angle=gen[:, 5:10], body_joints_angle=gen[:, 10:23],
              right_hand_joints_angle=gen[:, 23:28], left_hand_joints_angle=gen[:, 28:33],
              generator_global_rot_3d_pos_angle=gen[:, -3:], root_3d_pos=root)
    for i, name in enumerate(BONES):
        kw[name] = bone[:, i]
    return kw


@pytest.mark.parametrize("case,root_shape", [("gan133", None), ("video36", (4, 9, 3))])
def test_change_3d_joint_angle_like_the_generator_calls_it(golden, case, root_shape):
    from dhfk import camera
    g = golden(case)
    n = g["ang"].shape[0]
    gen = torch.zeros(n, 37, device="cuda")
    gen[:, :33] = T(g["ang"]); gen[:, 34:] = T(g["grot"])
    gen.requires_grad_(True)
    root = T(g["root"], True)
    root_in = root if root_shape is None else root.view(*root_shape)
    m = _model(batch=n)                                   # N is NOT pinned to args.batch_size any more
    w32 = m.change_3d_joint_angle(**_kwargs(gen, T(g["bone"]), root_in))
    assert w32.shape == (n, 32, 3) and w32.is_cuda
    assert_parity(w32.detach().cpu().numpy(), g["world32"], "world32")
    idx = [0, 1, 2, 3, 6, 7, 8, 12, 13, 15, 17, 18, 19, 25, 26, 27]
    fake = w32[:, idx].view(-1, 16, 3)                    # Fk_generator.py:259
    blk = g["cam_block"]
    cam = camera.GAN_torch_world_to_camera(fake, R=T(blk[0:4]).view(1, 4), t=T(blk[4:7]).view(1, 3))
    uv = camera.project_to_2d(cam, T(blk[7:16]).view(1, 9).repeat(n, 1))
    assert_parity(cam.detach().cpu().numpy(), g["cam"], "cam")
    assert_parity(uv.detach().cpu().numpy(), g["uv"], "uv")
    ((fake * T(g["g_world"])).sum() + (cam * T(g["g_cam"])).sum() + (uv * T(g["g_uv"])).sum()).backward()
    cond = projection_conditioning(g["cam"], g["world16"], g["cam_block"], g["g_uv"])
    gg = gen.grad.cpu().numpy()
    assert_parity(gg[:, :33], g["g_ang_wcu"], "g_ang", row_scale=cond)
    assert_parity(gg[:, 34:], g["g_grot_wcu"], "g_grot", row_scale=cond)
    assert_parity(root.grad.cpu().numpy(), g["g_root_wcu"], "g_root", row_scale=cond)
    # stateless: a second call with other lengths does not see the first one's (the reference mutates tables)
    w32b = m.change_3d_joint_angle(**_kwargs(gen.detach(), 2 * T(g["bone"]), root_in.detach()))
    assert not torch.allclose(w32b, w32.detach())


@pytest.mark.parametrize("n", [128, 96, 133, 31])
def test_wide_slot_tensor_path_full_and_ragged_tiles(golden, n):
    """The generator's [N,37] slot tensor goes to the kernels as one slab per tile and its gradient comes back as one
    [N,37] tensor (full tiles: slab store incl. the zero column 33; ragged tile: per-column stores into zeros).  The
    [N,32,3] result stays lazy: the generator's `[:, H36M_32_To_16_Table]` is the kernel's own output."""
    from dhfk import forward_kinematics_DH_model as fkm
    g = golden("gan133")
    gen = torch.full((n, 37), 7.0, device="cuda")          # column 33 holds junk the kernels must ignore
    gen[:, :33] = T(g["ang"][:n]); gen[:, 34:] = T(g["grot"][:n])
    gen.requires_grad_(True)
    slots = gen * 1.0                                      # a non-leaf, like the generator's range-mapped tensor
    root = T(g["root"][:n], True)
    kw = _kwargs(slots, T(g["bone"][:n]), root)
    kw["generator_global_rot_3d_pos_angle"] = slots[:, 34:37]
    w32 = _model(batch=n).change_3d_joint_angle(**kw)
    assert isinstance(w32, fkm.LazyWorld32) and w32._t32 is None
    fake = w32[:, [0, 1, 2, 3, 6, 7, 8, 12, 13, 15, 17, 18, 19, 25, 26, 27]]
    assert w32._t32 is None and fake.shape == (n, 16, 3)
    assert_parity(fake.detach().cpu().numpy(), g["world16"][:n], "world16")
    (fake * T(g["g_world"][:n])).sum().backward()
    gg = gen.grad.cpu().numpy()
    assert_parity(gg[:, :33], g["g_ang_w"][:n], "g_ang")
    assert_parity(gg[:, 34:], g["g_grot_w"][:n], "g_grot")
    assert np.all(gg[:, 33] == 0.0)
    assert_parity(root.grad.cpu().numpy(), g["g_root_w"][:n], "g_root")
    assert_parity(w32.tensor().detach().cpu().numpy(), g["world32"][:n], "world32 (materialised on demand)")


def test_init_fk_dh_angle_numpy_branch(golden):
    g = golden("kat")
    m = _model()
    out = m.init_Fk_DH_angle()
    assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == (32, 3)
    assert np.abs(out - g["tpose32"]).max() < 1e-6


def test_cpu_inputs_are_moved_like_the_reference_cuda_calls(golden):
    g = golden("gan133")
    n = 133
    gen = torch.zeros(n, 37); gen[:, :33] = torch.tensor(g["ang"])   # CPU tensors
    kw = _kwargs(gen, torch.tensor(g["bone"]), torch.tensor(g["root"]))
    kw["generator_global_rot_3d_pos_angle"] = torch.zeros(n, 3)      # whether_use_RT=False passes CPU zeros
    w32 = _model().change_3d_joint_angle(**kw)
    assert w32.is_cuda and w32.shape == (n, 32, 3)


def test_camera_functions_standalone(golden):
    from dhfk import camera
    g = golden("camera_ops")
    x = T(g["x"], True)
    uv = camera.project_to_2d(x, T(g["cam_rows9"]))
    assert_parity(uv.detach().cpu().numpy(), g["uv"], "uv")
    (uv * T(g["g_uv"])).sum().backward()
    cond = projection_conditioning(g["x"], g_uv=g["g_uv"])
    assert_parity(x.grad.cpu().numpy(), g["g_x"], "g_x", row_scale=cond)
    uv16 = camera.project_to_2d(T(g["x"]), T(g["cam_rows16"]))       # 16-column rows: only 9 are read
    assert torch.equal(uv16, uv.detach())
    xw = T(g["w_x"], True)
    cam = camera.GAN_torch_world_to_camera(xw, T(g["w_q"]).view(1, 4), T(g["w_t"]).view(1, 3))
    assert_parity(cam.detach().cpu().numpy(), g["w_cam"], "w_cam")
    (cam * T(g["w_g_cam"])).sum().backward()
    assert_parity(xw.grad.cpu().numpy(), g["w_g_x"], "w_g_x")


@pytest.mark.parametrize("mode", ["different", "same"])
def test_handler_but_generater_matches_reference(golden, mode):
    """The non-GAN sampler: same RNG draws as the reference, poses from ONE fused launch instead of 40
    numpy FK calls; the reference computes them in float64 and casts to float32."""
    from dhfk import Forward_Kinematics_DH_Model
    g = golden("sampler40")
    args = argparse.Namespace(batch_size=1, random_seed=5, single_or_multi_train_mode="single", architecture="3,3,3",
                              generator_whole_number=40, generator_choose_BoneLen=False,
                              generator_choose_root_pos=False, generator_global_rot=True, bone_len_scaler=mode)
    m = Forward_Kinematics_DH_Model(args, ["S1"], None)
    m.record_bone_len = [0.45, 0.45, 0.44, 0.44, 0.13, 0.13, 0.23, 0.26, 0.15, 0.15, 0.28, 0.28, 0.25, 0.25, 0.18]
    m.root_3d_pos = np.array([0.1, -0.2, 0.9])
    pos, ang, glob, bl, root = m.handler_but_generater()
    assert pos.shape == (40, 32, 3) and pos.dtype == np.float32
    assert np.array_equal(ang, g[mode + "_ang"])
    assert_parity(pos, g[mode + "_pos32"], "pos32")
    assert len(bl) == 40 and len(root) == 40 and len(glob) == 40


def test_camera_ops_other_joint_counts_take_the_point_kernels(golden):
    """16-joint poses run the tiled kernels; any other joint count (or an unaligned view) runs the one-thread-per-point
    kernels.  Both must give the reference's numbers: compare with the torch port (the reference's op sequence) on
    17- and 5-joint inputs and on an unaligned 16-joint view."""
    import torch_port
    from dhfk import camera
    g = golden("camera_ops")
    rng = np.random.RandomState(4)
    q, t = g["w_q"], g["w_t"]
    for joints in (17, 5):
        x = rng.uniform(-2, 2, (50, joints, 3)).astype(np.float32)
        x[..., 2] = rng.uniform(1.5, 6.0, (50, joints))
        rows = g["cam_rows9"][:50]
        xt, xc = T(x, True), torch.tensor(x, requires_grad=True)
        uv = camera.project_to_2d(xt, T(rows))
        ref = torch_port.project_to_2d(xc, torch.tensor(rows))
        assert_parity(uv.detach().cpu().numpy(), ref.detach().numpy(), "uv J=%d" % joints)
        gu = rng.randn(50, joints, 2).astype(np.float32)
        (uv * T(gu)).sum().backward(); (ref * torch.tensor(gu)).sum().backward()
        assert_parity(xt.grad.cpu().numpy(), xc.grad.numpy(), "g_x J=%d" % joints)
        wt, wc = T(x, True), torch.tensor(x, requires_grad=True)
        cam = camera.GAN_torch_world_to_camera(wt, T(q).view(1, 4), T(t).view(1, 3))
        refc = torch_port.world_to_camera(wc, torch.tensor(q).view(1, 4), torch.tensor(t).view(1, 3))
        assert_parity(cam.detach().cpu().numpy(), refc.detach().numpy(), "cam J=%d" % joints)
        gc = rng.randn(50, joints, 3).astype(np.float32)
        (cam * T(gc)).sum().backward(); (refc * torch.tensor(gc)).sum().backward()
        assert_parity(wt.grad.cpu().numpy(), wc.grad.numpy(), "g_world J=%d" % joints)
    # tiled and point kernels agree on the golden 16-joint case (bit for bit is not required, parity is)
    uv16 = camera.project_to_2d(T(g["x"]), T(g["cam_rows9"]))
    assert_parity(uv16.cpu().numpy(), g["uv"], "uv tiled")
    uv17 = camera.project_to_2d(torch.cat([T(g["x"]), T(g["x"][:, :1])], 1), T(g["cam_rows9"]))     # 17 joints: point kernel
    assert_parity(uv17[:, :16].cpu().numpy(), g["uv"], "uv point kernel")


@pytest.mark.parametrize("n", [1, 31, 33, 1000])
def test_scatter_32_layout_and_gradient(n):
    """dhfk_scatter32_*: the [N,32,3] slot layout and its adjoint are exact (copies and sums of copies), checked
    against the index arithmetic the reference's column writes amount to."""
    from dhfk.forward_kinematics_DH_model import scatter_16_to_32
    rng = np.random.RandomState(n)
    w = T(rng.randn(n, 16, 3), True)
    r = T(rng.randn(n, 3), True)
    out = scatter_16_to_32(w, r)
    idx = [0, 1, 2, 3, 6, 7, 8, 12, 13, 15, 17, 18, 19, 25, 26, 27]
    ref = r.detach().view(n, 1, 3).expand(n, 32, 3).clone()
    ref[:, idx] = w.detach()
    ref[:, 14] = w.detach()[:, 9]
    assert torch.equal(out, ref)
    g = T(rng.randn(n, 32, 3))
    (out * g).sum().backward()
    g16 = g[:, idx].clone()
    g16[:, 9] += g[:, 14]
    free = [s for s in range(32) if s not in idx + [14]]
    assert len(free) == 15
    assert torch.allclose(w.grad, g16, rtol=0, atol=1e-6) and torch.equal(w.grad[:, :9], g16[:, :9])
    assert torch.allclose(r.grad, g[:, free].sum(1), rtol=1e-6, atol=1e-6)


def test_scatter_16_to_32_reproduces_the_reference_tensor(golden):
    """Against the reference's own [N,32,3] output (goldens): pure data movement, bit-exact; root given as [B,F,3] too."""
    from dhfk.forward_kinematics_DH_model import scatter_16_to_32
    g = golden("gan133")
    out = scatter_16_to_32(T(g["world16"]), T(g["root"]))
    assert np.array_equal(out.cpu().numpy(), g["world32"])
    gv = golden("video36")
    out = scatter_16_to_32(T(gv["world16"]), T(gv["root"]).view(4, 9, 3))
    assert np.array_equal(out.cpu().numpy(), gv["world32"])


def test_cpu_float64_callers_get_their_own_device_and_dtype_back(golden):
    """common/camera.py::wrap feeds CPU float64 tensors and calls .numpy() on the result (the
    `--data_enhancement_method normal` path, model_fk_gan_train.py:74): the kernels run on the GPU in fp32, the result
    comes back where and as what the caller's tensors were."""
    from dhfk import camera
    g = golden("camera_ops")
    x = torch.tensor(g["x"], dtype=torch.float64)
    uv = camera.project_to_2d(x, torch.tensor(g["cam_rows9"], dtype=torch.float64))
    assert uv.device.type == "cpu" and uv.dtype == torch.float64
    assert_parity(uv.numpy(), g["uv"], "uv")
    cam = camera.GAN_torch_world_to_camera(torch.tensor(g["w_x"], dtype=torch.float64),
                                           torch.tensor(g["w_q"], dtype=torch.float64).view(1, 4),
                                           torch.tensor(g["w_t"], dtype=torch.float64).view(1, 3))
    assert cam.device.type == "cpu" and cam.dtype == torch.float64
    assert_parity(cam.numpy(), g["w_cam"], "w_cam")
