"""Host-side logic that needs no GPU: sharding, synthetic inputs, the reference-shaped class's
bookkeeping, the 16->32 scatter, drop-in patching, and the loud failure without CUDA."""
import sys
import types

import numpy as np
import pytest
import torch

from dhfk import parallel, synthetic, tables
from dhfk.forward_kinematics_DH_model import Forward_Kinematics_DH_Model, scatter_16_to_32


def test_shard_rows_partition():
    for n in (0, 1, 7, 96, 1000, 16777216):
        for r in (1, 2, 3, 4, 8):
            spans = [parallel.shard_rows(n, k, r) for k in range(r)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(r - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_rows(10, 2, 2)


def test_synthetic_inputs_follow_generator_ranges():
    d = synthetic.gan_like(4096, seed=1)
    assert d["ang"].dtype == np.float32 and d["ang"].shape == (4096, 33)
    lo, hi = tables.GAN_ANGLE_RANGE[:33, 0], tables.GAN_ANGLE_RANGE[:33, 1]
    assert (d["ang"] >= lo - 1e-3).all() and (d["ang"] <= hi + 1e-3).all()
    assert (d["ang"][:, [4, 9, 22, 23, 28]] == 0).all()
    assert np.abs(d["grot"]).max() <= 180 and d["grot"].std() > 90
    ratio = d["bone"][:, 7:8] / tables.BONE_TEMPLATES[:, 7][None, :]
    assert (np.abs(ratio - 1).min(1) < 1e-6).all()          # thorax is never scaled
    assert (d["bone"] > 0.08).all() and (d["bone"] < 0.6).all()
    assert np.array_equal(d["bone"][:, 0] / d["bone"][:, 1] > 0.99, np.ones(4096, bool))  # symmetric groups
    d2 = synthetic.gan_like(4096, seed=1)
    assert all(np.array_equal(d[k], d2[k]) for k in d)       # deterministic
    s = synthetic.gan_like(512, seed=2, root_mode="generator", angle_mode="stress")
    assert np.abs(s["root"]).max() <= 10 and np.abs(s["ang"]).max() <= 180


def test_reference_shaped_class_without_gpu():
    import argparse
    args = argparse.Namespace(batch_size=8, random_seed=3, single_or_multi_train_mode="multi", architecture="3,3")
    m = Forward_Kinematics_DH_Model(args, ["S1", "S5"], None)
    assert m.real_used_num == 9 and m.GAN_BATCH_SIZE == 8
    assert m.random.randint(0, 1000) == np.random.RandomState(3).randint(0, 1000)
    assert m.record_bone_len == [] and list(m.root_3d_pos) == [0, 0, 0]
    assert m.body_joints_alpha == tables.ALPHA_DEG[10:23].tolist()
    r = np.random.RandomState(5)
    m.set_random_state(r)
    assert m.random_state() is r
    # dataset bookkeeping draws follow the reference order (subject, action, camera, frame)
    m.dataSet_world_3d_pos = {"S1": {"a": {0: np.arange(48 * 3, dtype=np.float64).reshape(3, 16, 3)}},
                              "S5": {"b": {0: np.ones((2, 16, 3))}}}
    m.dataSet_2d_pos = m.dataSet_world_3d_pos
    m.set_random_state(np.random.RandomState(0))
    m.get_bone_len_from_dataSet()
    assert len(m.record_bone_len) == 15
    m.get_root_3d_pos_from_dataSet()
    assert m.root_3d_pos.shape == (3,)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.init_Fk_DH_angle()
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            from dhfk import fk_world16
            fk_world16(torch.zeros(2, 33), torch.zeros(2, 3), torch.ones(2, 15), torch.zeros(2, 3))


def test_dropin_install_patches_imported_reference_modules(monkeypatch):
    from dhfk import camera, dropin
    fk = types.ModuleType("models_Fk_GAN.forward_kinematics_DH_model")
    fk.Forward_Kinematics_DH_Model = object
    cam = types.ModuleType("common.camera")
    cam.GAN_torch_world_to_camera = cam.project_to_2d = cam.normalize_screen_coordinates = (lambda *a: None)
    keep = cam.normalize_screen_coordinates
    train = types.ModuleType("models_Fk_GAN.model_fk_gan_train")
    train.project_to_2d = train.GAN_torch_world_to_camera = train.Forward_Kinematics_DH_Model = None
    monkeypatch.setitem(sys.modules, "models_Fk_GAN.forward_kinematics_DH_model", fk)
    monkeypatch.setitem(sys.modules, "common.camera", cam)
    monkeypatch.setitem(sys.modules, "models_Fk_GAN.model_fk_gan_train", train)
    patched = dropin.install()
    assert fk.Forward_Kinematics_DH_Model is Forward_Kinematics_DH_Model
    assert cam.project_to_2d is camera.project_to_2d and cam.GAN_torch_world_to_camera is camera.GAN_torch_world_to_camera
    assert cam.normalize_screen_coordinates is keep
    assert train.project_to_2d is camera.project_to_2d and train.Forward_Kinematics_DH_Model is Forward_Kinematics_DH_Model
    assert len(patched) >= 5


def test_camera_dropin_assertions_match_reference():
    from dhfk import camera
    with pytest.raises(AssertionError):
        camera.project_to_2d(torch.zeros(4, 16, 2), torch.zeros(4, 9))       # last dim must be 3
    with pytest.raises(AssertionError):
        camera.project_to_2d(torch.zeros(4, 16, 3), torch.zeros(9))          # camera_params.ndim == 2
    with pytest.raises(AssertionError):
        camera.project_to_2d(torch.zeros(4, 16, 3), torch.zeros(4, 10))      # 9 or 16 columns
    with pytest.raises(AssertionError):
        camera.project_to_2d(torch.zeros(4, 16, 3), torch.zeros(5, 9))       # batch mismatch
    with pytest.raises(AssertionError):
        camera.GAN_torch_world_to_camera(torch.zeros(4, 16, 3), torch.zeros(1, 3), torch.zeros(1, 3))


def _sampler_args(mode):
    import argparse
    return argparse.Namespace(batch_size=1, random_seed=5, single_or_multi_train_mode="single", architecture="3,3,3",
                              generator_whole_number=40, generator_choose_BoneLen=False,
                              generator_choose_root_pos=False, generator_global_rot=True, bone_len_scaler=mode)


@pytest.mark.parametrize("mode", ["different", "same"])
def test_sampler_draws_match_reference_rng_order(golden, mode):
    """handler_but_generater's host-side draws (forward_kinematics_DH_model.py:931-1113) must consume the
    RandomState exactly like the reference so that seeded runs produce the same poses."""
    g = golden("sampler40")
    m = Forward_Kinematics_DH_Model(_sampler_args(mode), ["S1"], None)
    m.record_bone_len = [0.45, 0.45, 0.44, 0.44, 0.13, 0.13, 0.23, 0.26, 0.15, 0.15, 0.28, 0.28, 0.25, 0.25, 0.18]
    m.root_3d_pos = np.array([0.1, -0.2, 0.9])
    ang, glob, bone, root = m.sample_normal_mode()
    assert np.array_equal(ang, g[mode + "_ang"])            # bit-exact: same draws, same clipping
    assert np.array_equal(glob, g[mode + "_glob"])
    assert m.random.randint(0, 1 << 30) == int(g[mode + "_next_draw"][0])   # RNG left in the same state
    assert bone.shape == (40, 15) and np.allclose(bone[:, 7], 0.26)          # thorax never scaled
    assert (ang[0] == 0).all() and (glob[0] == 0).all()                      # frame 0 is the rest pose


def test_scalar_dh_matrix_and_rotation_helpers(c_oracle):
    import ctypes
    from dhfk.forward_kinematics_DH_model import dh_matrix, rotationMatrix
    lib = c_oracle.lib()
    T = np.zeros(16); R = np.zeros(9)
    d = ctypes.c_double
    lib.dhfk_oracle_dh_matrix(d(-90.0), d(0.3), d(0.2), d(37.0), T.ctypes.data_as(ctypes.POINTER(d)))
    assert np.allclose(dh_matrix(-90.0, 0.3, 0.2, 37.0, None), T.reshape(4, 4), atol=1e-15)
    lib.dhfk_oracle_rotation_matrix(d(10.0), d(20.0), d(30.0), R.ctypes.data_as(ctypes.POINTER(d)))
    assert np.allclose(rotationMatrix(10.0, 20.0, 30.0, None), R.reshape(3, 3), atol=1e-15)
    # tensor inputs: one matrix per row, the same numbers (pinned to the reference itself in test_dropin_reference.py)
    Tt = dh_matrix(torch.full((3,), -90.0), torch.full((3,), 0.3), torch.full((3,), 0.2), torch.full((3,), 37.0), None)
    assert Tt.shape == (3, 4, 4) and np.allclose(Tt[1].numpy(), T.reshape(4, 4), atol=1e-6)
    Rt = rotationMatrix(torch.full((3,), 10.0), torch.full((3,), 20.0), torch.full((3,), 30.0), None)
    assert Rt.shape == (3, 3, 3) and np.allclose(Rt[2].numpy(), R.reshape(3, 3), atol=1e-6)


def test_dropin_install_widening_flags(monkeypatch):
    """install(generators/critics/loader_refresh=True) rebinds exactly the symbols INTEGRATION.md lists and nothing
    else; without the flags those symbols are left alone."""
    from dhfk import Fk_discriminator, Fk_generator, dataloader_update, dropin
    sentinel = object()
    mods = {}
    for name, syms in (("models_Fk_GAN.Fk_discriminator", ("special_KCS_Input_transform", "video_mode_special_KCS_Input_transform",
                                                           "Fk_3D_Discriminator", "calc_gradient_penalty")),
                       ("models_Fk_GAN.Fk_generator", ("Fk_Generator", "Video_Fk_Generator", "GAN_angle_range_table")),
                       ("function_aug.dataloader_update", ("random_bl_aug", "dataloader_update", "project_to_2d")),
                       ("models_Fk_GAN.video_mode_operate", ("random_bl_aug", "video_mode_random_bl_aug",
                                                             "video_mode_dataloader_update")),
                       ("run_Fk_GAN", ("dataloader_update",))):
        m = types.ModuleType(name)
        for s in syms:
            setattr(m, s, sentinel)
        monkeypatch.setitem(sys.modules, name, m)
        mods[name] = m
    dropin.install()
    assert mods["models_Fk_GAN.Fk_discriminator"].special_KCS_Input_transform is sentinel
    assert mods["models_Fk_GAN.Fk_generator"].Fk_Generator is sentinel
    assert mods["function_aug.dataloader_update"].random_bl_aug is sentinel
    patched = dropin.install(generators=True, critics=True, loader_refresh=True)
    d, g = mods["models_Fk_GAN.Fk_discriminator"], mods["models_Fk_GAN.Fk_generator"]
    assert d.special_KCS_Input_transform is Fk_discriminator.special_KCS_Input_transform
    assert d.video_mode_special_KCS_Input_transform is Fk_discriminator.video_mode_special_KCS_Input_transform
    assert d.Fk_3D_Discriminator is sentinel and d.calc_gradient_penalty is sentinel      # critic classes stay the reference's
    assert g.Fk_Generator is Fk_generator.Fk_Generator and g.Video_Fk_Generator is Fk_generator.Video_Fk_Generator
    assert g.GAN_angle_range_table is sentinel
    du, vo = mods["function_aug.dataloader_update"], mods["models_Fk_GAN.video_mode_operate"]
    assert du.random_bl_aug is dataloader_update.random_bl_aug and du.dataloader_update is dataloader_update.dataloader_update
    assert vo.random_bl_aug is dataloader_update.random_bl_aug
    assert vo.video_mode_random_bl_aug is dataloader_update.video_mode_random_bl_aug
    assert vo.video_mode_dataloader_update is dataloader_update.video_mode_dataloader_update
    assert mods["run_Fk_GAN"].dataloader_update is dataloader_update.dataloader_update
    assert any("special_KCS_Input_transform" in p for p in patched)


def test_bone_length_templates_table_matches_reference_file(golden):
    """The in-package copy of data_extra/bone_length_npy/hm36s15678_bl_templates.npy is bit-identical to the file."""
    from dhfk import dataloader_update, tables
    g = golden("tables")
    assert np.array_equal(tables.BONE_TEMPLATES_GANUTILS_ORDER, g["bone_templates"].astype(np.float32))
    assert np.array_equal(dataloader_update.bone_length_templates(), g["bone_templates"].astype(np.float32))


def test_shared_angle_view_detection():
    """The five angle kwargs as column slices of one [N,37] tensor (how the generator passes them) are recognised and
    replaced by ONE strided view of their base -- in kernel order, no torch.cat; anything else is left to torch.cat."""
    from dhfk.forward_kinematics_DH_model import _shared_angle_view as view
    g = torch.randn(7, 37, requires_grad=True) * 1.0
    sl = lambda a, b: g[:, a:b]
    v = view(sl(0, 5), sl(5, 10), sl(10, 23), sl(23, 28), sl(28, 33))
    assert v is not None and v.shape == (7, 33) and v.data_ptr() == g.data_ptr() and v.stride() == (37, 1)
    assert v._base is g                                             # autograd flows to the generator's tensor
    g2 = torch.randn(4, 40)
    v2 = view(g2[:, 2:7], g2[:, 7:12], g2[:, 12:25], g2[:, 25:30], g2[:, 30:35])
    assert v2.shape == (4, 33) and v2.storage_offset() == 2
    assert view(sl(5, 10), sl(0, 5), sl(10, 23), sl(23, 28), sl(28, 33)) is None          # not in generator order
    assert view(g2[:4, 0:5], sl(5, 10), sl(10, 23), sl(23, 28), sl(28, 33)) is None       # different bases
    assert view(*(torch.randn(7, w) for w in (5, 5, 13, 5, 5))) is None                   # independent tensors
    assert view(sl(0, 5), sl(5, 10), sl(10, 23), sl(23, 28), g[:, 28:33][::2]) is None    # a row-strided slice


def test_lazy_world32_answers_the_joint_gather_without_materialising(monkeypatch):
    """change_3d_joint_angle's [N,32,3] result (lazy 32-slot scatter): `x[:, H36M_32_To_16_Table]`
    (Fk_generator.py:259) must hand back the kernel's [N,16,3] output itself; anything else gets the real tensor."""
    from dhfk import forward_kinematics_DH_model as fkm
    calls = []

    def fake_scatter(w16, root):                      # the layout dhfk_scatter32_forward produces, in plain torch
        calls.append(1)
        out = root.reshape(-1, 1, 3).expand(-1, 32, 3).clone()
        out[:, tables.H36M_32_To_16_Table] = w16
        out[:, 14] = w16[:, 9]
        return out

    monkeypatch.setattr(fkm, "_scatter32", fake_scatter)
    w16 = torch.randn(5, 16, 3, requires_grad=True)
    root = torch.randn(5, 3)
    x = fkm.LazyWorld32(w16 * 1.0, root)
    assert x.shape == (5, 32, 3) and x.size(1) == 32 and x.dim() == 3 and len(x) == 5 and x.requires_grad
    got = x[:, tables.H36M_32_To_16_Table]
    assert got is x._w16 and not calls                # the generator's index: no scatter, autograd history intact
    assert x[:, np.array(tables.H36M_32_To_16_Table)] is x._w16 and x[:, torch.tensor(tables.H36M_32_To_16_Table)] is x._w16
    got.view(-1, 48).sum().backward()
    assert torch.equal(w16.grad, torch.ones_like(w16))
    # every other use sees the real [N,32,3] tensor, built once
    assert torch.equal(x[:, 14], x._w16[:, 9]) and len(calls) == 1
    assert torch.equal(x[:, [0, 1, 2]], x.tensor()[:, [0, 1, 2]])
    assert torch.equal(x[2], x.tensor()[2])
    assert torch.equal(torch.sum(x, dim=1), x.tensor().sum(dim=1))            # __torch_function__
    assert torch.equal(torch.cat([x, x], dim=0), torch.cat([x.tensor(), x.tensor()], dim=0))
    assert torch.equal(x.detach().view(5, 96), x.tensor().detach().view(5, 96))   # attribute / method fall-through
    assert torch.equal(x * 2.0 - x, x.tensor()) and torch.equal(-x, -x.tensor())
    assert np.array_equal(np.asarray(x), x.tensor().detach().numpy())
    assert len(calls) == 1
    assert torch.equal(x[:, 4], root) and torch.equal(x[:, 31], root)          # free slots hold the root


def test_peer_exchange_range_checks_without_gpu():
    """dhfk.parallel.PeerExchange.allreduce validates the element range before anything reaches the library."""
    import pytest
    import torch
    from dhfk import parallel
    buf = torch.zeros(64)
    px = parallel.PeerExchange(buf, torch.zeros(8, dtype=torch.int32), torch.zeros(1, dtype=torch.int32),
                               [buf.data_ptr(), buf.data_ptr()], [0x1000, 0x2000], 0, 0, 2, 16, 512, 100, None)
    assert px.multicast is False and px.peer_min_ctas == 32
    assert px.allreduce(8, 8) == 0                      # empty range: nothing launched
    for lo, hi in ((2, 8), (0, 6), (-4, 8), (0, 68), (16, 8)):
        with pytest.raises(ValueError):
            px.allreduce(lo, hi)
    # off NCCL there is nothing to set up, and FlatGradBuffer keeps an ordinary tensor
    assert parallel.PeerExchange.create(64, torch.device("cpu")) is None
    lin = torch.nn.Linear(3, 2)
    fb = parallel.FlatGradBuffer(lin.parameters(), peer_exchange=True)
    assert fb.peer is None and fb.flat.numel() == 8 + 4 and lin.weight.grad.data_ptr() == fb.flat.data_ptr()
