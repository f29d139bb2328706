"""Build-container only (skipped where /root/reference is absent, e.g. on the GPU box): dropin.install() against
the REAL reference modules -- every symbol INTEGRATION.md lists exists there under that name, gets rebound, and the
reference's own classes then route into the native entry points (which, without a GPU, raise instead of falling
back to any CPU path)."""
import argparse
import importlib
import sys

import pytest
import torch

import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not mounted")

REF_MODULES = ("models_Fk_GAN.model_fk_gan_train", "models_Fk_GAN.video_GAN_fun", "function_aug.dataloader_update",
               "models_Fk_GAN.video_mode_operate", "models_Fk_GAN.Fk_discriminator", "models_Fk_GAN.Fk_generator",
               "common.camera", "models_Fk_GAN.forward_kinematics_DH_model")


@pytest.fixture()
def reference_modules():
    rh.import_reference()
    for name in ("progress", "progress.bar"):
        if name not in sys.modules:
            rh._stub_module(name)
    mods = {m: importlib.import_module(m) for m in REF_MODULES}
    saved = {m: dict(vars(mod)) for m, mod in mods.items()}
    dis = mods["models_Fk_GAN.Fk_discriminator"]
    forwards = {c: getattr(dis, c).forward for c in ("Video_motion_Fk_3D_Discriminator", "Video_motion_Fk_2D_Discriminator")}
    yield mods
    for m, mod in mods.items():          # undo the patching: other tests use the unmodified reference
        for k, v in saved[m].items():
            setattr(mod, k, v)
    for c, f in forwards.items():        # install(critics=True) swaps these two methods on the reference's own classes
        getattr(dis, c).forward = f


def test_install_rebinds_the_real_reference_symbols(reference_modules):
    from dhfk import Fk_discriminator, Fk_generator, camera, dataloader_update, dropin
    from dhfk import Forward_Kinematics_DH_Model
    m = reference_modules
    before = m["models_Fk_GAN.Fk_discriminator"].special_KCS_Input_transform
    patched = dropin.install(generators=True, critics=True, loader_refresh=True)
    assert m["models_Fk_GAN.forward_kinematics_DH_model"].Forward_Kinematics_DH_Model is Forward_Kinematics_DH_Model
    for name in ("models_Fk_GAN.model_fk_gan_train", "models_Fk_GAN.video_GAN_fun"):
        assert m[name].GAN_torch_world_to_camera is camera.GAN_torch_world_to_camera
        assert m[name].project_to_2d is camera.project_to_2d
        assert m[name].Fk_Generator is Fk_generator.Fk_Generator
        assert m[name].Video_Fk_Generator is Fk_generator.Video_Fk_Generator
    assert m["common.camera"].project_to_2d is camera.project_to_2d
    assert m["function_aug.dataloader_update"].project_to_2d is camera.project_to_2d
    assert m["function_aug.dataloader_update"].random_bl_aug is dataloader_update.random_bl_aug
    assert m["function_aug.dataloader_update"].dataloader_update is dataloader_update.dataloader_update
    assert m["models_Fk_GAN.video_mode_operate"].video_mode_random_bl_aug is dataloader_update.video_mode_random_bl_aug
    assert (m["models_Fk_GAN.video_mode_operate"].video_mode_dataloader_update
            is dataloader_update.video_mode_dataloader_update)
    # what the native video refresh borrows from the reference at call time exists under those names
    assert callable(m["models_Fk_GAN.video_mode_operate"].GAN_video_ChunkedGenerator)
    assert m["models_Fk_GAN.video_mode_operate"].video_receptive_field([3, 3, 3]) == 27
    d = m["models_Fk_GAN.Fk_discriminator"]
    assert d.special_KCS_Input_transform is Fk_discriminator.special_KCS_Input_transform is not before
    assert d.video_mode_special_KCS_Input_transform is Fk_discriminator.video_mode_special_KCS_Input_transform
    # the two motion critics keep their class (constructor, sub-modules, state dict) and get the fused forward
    assert d.Video_motion_Fk_3D_Discriminator.forward is Fk_discriminator.video_motion_3d_forward
    assert d.Video_motion_Fk_2D_Discriminator.forward is Fk_discriminator.video_motion_2d_forward
    ref3 = d.Video_motion_Fk_3D_Discriminator("cpu", argparse.Namespace(
        video_Dis_DenseDim_3D=8, motion_Dis_whether_use_3dPos_branch=True, motion_Dis_whether_use_3dDiff_branch=True), 9)
    ours3 = Fk_discriminator.Video_motion_Fk_3D_Discriminator("cpu", ref3.args, 9)
    assert {k: tuple(v.shape) for k, v in ref3.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in ours3.state_dict().items()}
    ref2 = d.Video_motion_Fk_2D_Discriminator("cpu", argparse.Namespace(video_Dis_DenseDim_2D=8), 9)
    ours2 = Fk_discriminator.Video_motion_Fk_2D_Discriminator("cpu", ref2.args, 9)
    assert {k: tuple(v.shape) for k, v in ref2.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in ours2.state_dict().items()}
    refd = d.Fk_3D_Discriminator("cpu", argparse.Namespace(Dis_DenseDim_3D=8))
    assert {k: tuple(v.shape) for k, v in refd.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in Fk_discriminator.Fk_3D_Discriminator("cpu", refd.args).state_dict().items()}
    ref2d = d.Fk_2D_Discriminator(argparse.Namespace(Dis_DenseDim_2D=8))
    assert {k: tuple(v.shape) for k, v in ref2d.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in Fk_discriminator.Fk_2D_Discriminator(ref2d.args).state_dict().items()}
    assert len(patched) >= 22


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_reference_critic_routes_into_the_native_transform(reference_modules):
    """The reference's Fk_3D_Discriminator looks special_KCS_Input_transform up as a module global
    (Fk_discriminator.py:190): after install(critics=True) its forward reaches the native entry point, which
    refuses to run without a CUDA device."""
    from dhfk import dropin
    d = reference_modules["models_Fk_GAN.Fk_discriminator"]
    D = d.Fk_3D_Discriminator(torch.device("cpu"), argparse.Namespace(Dis_DenseDim_3D=16))
    x = torch.randn(4, 16, 3)
    assert D(x).shape == (4, 1)                                   # the unmodified reference runs on the CPU
    dropin.install(critics=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        D(x)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_reference_generator_class_is_state_dict_compatible(reference_modules):
    """Same constructor, sub-module names and parameter shapes as the reference generator (Fk_generator.py:80-112),
    so checkpoints move both ways."""
    from dhfk import Fk_generator as native
    g = reference_modules["models_Fk_GAN.Fk_generator"]
    fkmod = reference_modules["models_Fk_GAN.forward_kinematics_DH_model"]
    args = argparse.Namespace(batch_size=8, random_seed=0, single_or_multi_train_mode="single", architecture="3,3,3",
                              GAN_OUTPUT_DIM=35, Gen_DenseDim=32, GAN_whether_use_preAngle=True, whether_use_RT=True,
                              bone_len_scaler="different", record_all_picture=False, checkpoint="/tmp")
    ref_fk = fkmod.Forward_Kinematics_DH_Model(args, ["S1"], None)
    G_ref = g.Fk_Generator(ref_fk, args, torch.device("cpu"))
    from dhfk import Forward_Kinematics_DH_Model
    G_nat = native.Fk_Generator(Forward_Kinematics_DH_Model(args, ["S1"], None), args, torch.device("cpu"))
    sd_ref, sd_nat = G_ref.state_dict(), G_nat.state_dict()
    assert list(sd_ref) == list(sd_nat)
    assert all(sd_ref[k].shape == sd_nat[k].shape for k in sd_ref)
    G_nat.load_state_dict(sd_ref)


def test_tensor_dh_matrix_and_rotation_matrix_match_the_reference():
    """forward_kinematics_DH_model.py:80-116 / :141-191 on tensor inputs (SURVEY 8b lists both as module globals to
    preserve): same matrices as the unmodified reference, and differentiable."""
    ref = rh.import_reference()
    from dhfk import forward_kinematics_DH_model as ours
    n = 24
    args = rh.make_args(n)
    g = torch.Generator().manual_seed(2)
    theta = (torch.rand(n, generator=g) * 720 - 360).requires_grad_(True)
    a, d = torch.rand(n, generator=g), torch.rand(n, generator=g)
    for alpha_deg in (0.0, 90.0, -90.0, 37.5):
        alpha = torch.full((n,), alpha_deg)
        want = ref.fk.dh_matrix(alpha, a, d, theta, args)
        got = ours.dh_matrix(alpha, a, d, theta, args)
        assert got.shape == (n, 4, 4) and torch.allclose(got, want, atol=1e-6, rtol=0)
        (gw,) = torch.autograd.grad((want * torch.arange(16.0).view(4, 4)).sum(), theta)
        (gg,) = torch.autograd.grad((got * torch.arange(16.0).view(4, 4)).sum(), theta)
        assert torch.allclose(gg, gw, atol=1e-5, rtol=1e-5)
    ax, ay, az = (torch.rand(n, generator=g) * 360 - 180 for _ in range(3))
    assert torch.allclose(ours.rotationMatrix(ax, ay, az, args), ref.fk.rotationMatrix(ax, ay, az, args), atol=1e-6, rtol=0)
    import numpy as np
    assert np.allclose(ours.dh_matrix(-90.0, 0.3, 0.2, 33.0), ref.fk.dh_matrix(-90.0, 0.3, 0.2, 33.0, args), atol=1e-12)
    assert np.allclose(ours.rotationMatrix(10.0, -20.0, 30.0), ref.fk.rotationMatrix(10.0, -20.0, 30.0, args), atol=1e-12)
