"""Generate tests/golden/*.npz by running the UNMODIFIED reference (torch fp32, CPU branch).

Run in the build container only (needs /root/reference):
    python oracle/make_golden.py
The fixtures are committed; nothing that runs on the GPU box reads /root/reference.

Fixtures
  tables.npz       constant tables read off the reference modules/objects + the structural
                   dependency matrix of its autograd Jacobian (which angle moves which output)
  kat.npz          KAT-1 T-pose (init_Fk_DH_angle, numpy float64 branch) and KAT-2 bent pose
  gan133.npz       133 generator-like poses (ragged vs the 32-row kernel tile), S1/cam0, in-volume root
  stress200.npz    200 poses, U(-180,180) on every slot, root 10*tanh(randn) (clamp active), S7/cam2
  video36.npz      multi-frame mode: B=4 clips x F=9 frames, root given as [B,F,3]
  camera_ops.npz   GAN_torch_world_to_camera / project_to_2d on their own, per-row intrinsics (9 and 16 cols)
  sampler40.npz    handler_but_generater (non-GAN sampler) draws + poses for a fixed seed
  generator.npz    Fk_Generator / Video_Fk_Generator epilogue (tanh, slot scatter, range map, scaler) + FK,
                   outputs and d/d(raw network output), for a known raw last-layer output
  retarget.npz     random_bl_aug / video_mode_random_bl_aug + per-row project_to_2d, seeded
                   np.random so the drawn template rows are part of the fixture
  critic.npz       critic input transforms: special_KCS_Input_transform (30 cols) and its video variant
                   (15 cols), the train loop's root-centring and left/right flip, autograd gradients, and the
                   reference Fk_3D_Discriminator's WGAN-GP gradient penalty + parameter gradients (double backward
                   through the KCS transform) for a seeded state dict
  video_critic.npz the motion critics of the multi-frame mode (video part): the unmodified
                   Video_motion_Fk_3D_Discriminator / Video_motion_Fk_2D_Discriminator (F = 9, seeded state dicts) on
                   a clip batch and on its torch.flip(dims=[1]) playback reverse -- the feature tensors each branch
                   receives (captured with forward pre-hooks), outputs, input gradients, WGAN-GP penalty and its
                   parameter gradients
  gan_loop.npz     the reference's own training loops, unpatched, on CPU (oracle/ref_loop.py): 6 iterations of
                   GAN_solutions_FK_generator (batch 32) and of video_mode_GAN_solutions_FK_generator (8 clips x 9
                   frames) on seeded synthetic loaders -- every scalar the loop reports (per critic step), the generator
                   gradients at the generator step, the fake-pair buffer it leaves behind

    python oracle/make_golden.py retarget      # regenerate only the named fixtures
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def t(x, grad=False):
    return torch.tensor(np.asarray(x, dtype=np.float32), requires_grad=grad)


def run_case(inp, cam_blk, grads, mode="single", architecture="3,3,3", root_shape=None):
    ref = rh.import_reference()
    ang, grot, root = t(inp["ang"], True), t(inp["grot"], True), t(inp["root"], True)
    bone = t(inp["bone"])
    n = ang.shape[0]
    frames = 1
    if mode == "multi":
        for f in architecture.split(","):
            frames *= int(f)
    model = ref.fk.Forward_Kinematics_DH_Model(rh.make_args(n // frames, mode, architecture), ["S1"], None)
    root_in = root if root_shape is None else root.view(*root_shape)
    kw = dict(right_leg_joints_angle=ang[:, 0:5], left_leg_joints_angle=ang[:, 5:10],
              body_joints_angle=ang[:, 10:23], right_hand_joints_angle=ang[:, 23:28],
              left_hand_joints_angle=ang[:, 28:33], generator_global_rot_3d_pos_angle=grot, root_3d_pos=root_in)
    for i, name in enumerate(rh.BONE_KWARGS):
        kw[name] = bone[:, i]
    w32 = model.change_3d_joint_angle(**kw)
    w16 = w32[:, ref.h36m.H36M_32_To_16_Table]
    q, tt = t(cam_blk[0:4]).view(1, 4), t(cam_blk[4:7]).view(1, 3)
    cp = t(cam_blk[7:16]).view(1, 9).repeat(n, 1)
    cam = ref.camera.GAN_torch_world_to_camera(w16, R=q, t=tt)
    uv = ref.camera.project_to_2d(cam, cp)
    out = dict(inp)
    out.update(cam_block=cam_blk, world32=w32.detach().numpy(), world16=w16.detach().numpy(),
               cam=cam.detach().numpy(), uv=uv.detach().numpy())
    # gradient sets: (world only) and (world + cam + uv) and (world + uv)
    gw, gc, gu = t(grads["g_world"]), t(grads["g_cam"]), t(grads["g_uv"])
    out.update(g_world=grads["g_world"], g_cam=grads["g_cam"], g_uv=grads["g_uv"])
    for tag, loss in (("w", (w16 * gw).sum()),
                      ("wu", (w16 * gw).sum() + (uv * gu).sum()),
                      ("wcu", (w16 * gw).sum() + (cam * gc).sum() + (uv * gu).sum())):
        ga, gg, gr = torch.autograd.grad(loss, (ang, grot, root), retain_graph=True)
        out["g_ang_" + tag] = ga.numpy()
        out["g_grot_" + tag] = gg.numpy()
        out["g_root_" + tag] = gr.numpy()
    return out


def tables_fixture():
    ref = rh.import_reference()
    m = ref.fk.Forward_Kinematics_DH_Model(rh.make_args(1), ["S1"], None)
    alpha = np.array(m.right_leg_joints_alpha + m.left_leg_joints_alpha + m.body_joints_alpha +
                     m.right_hand_joints_alpha + m.left_hand_joints_alpha, dtype=np.float32)
    theta0 = np.array(m.right_leg_joints_theta + m.left_leg_joints_theta + m.body_joints_theta +
                      m.right_hand_joints_theta + m.left_hand_joints_theta, dtype=np.float32)
    # bone-length slots: call the reference with distinct lengths and read its a/d tables back
    lens = np.array([0.5 + 0.01 * i for i in range(15)], dtype=np.float32)
    kw = dict(right_leg_joints_angle=torch.zeros(1, 5), left_leg_joints_angle=torch.zeros(1, 5),
              body_joints_angle=torch.zeros(1, 13), right_hand_joints_angle=torch.zeros(1, 5),
              left_hand_joints_angle=torch.zeros(1, 5), generator_global_rot_3d_pos_angle=torch.zeros(1, 3),
              root_3d_pos=torch.zeros(1, 3))
    for i, name in enumerate(rh.BONE_KWARGS):
        kw[name] = torch.tensor([lens[i]])
    m.change_3d_joint_angle(**kw)
    a = torch.cat([m.GAN_right_leg_joints_a, m.GAN_left_leg_joints_a, m.GAN_body_joints_a,
                   m.GAN_right_hand_joints_a, m.GAN_left_hand_joints_a], 1)[0].numpy()
    d = torch.cat([m.GAN_right_leg_joints_d, m.GAN_left_leg_joints_d, m.GAN_body_joints_d,
                   m.GAN_right_hand_joints_d, m.GAN_left_hand_joints_d], 1)[0].numpy()
    kind = np.zeros(33, np.int32); bone = -np.ones(33, np.int32); sign = np.zeros(33, np.int32)
    for j in range(33):
        for arr, k in ((a, 1), (d, 2)):
            hit = [i for i in range(15) if abs(abs(arr[j]) - lens[i]) < 1e-6]
            if hit:
                kind[j], bone[j], sign[j] = k, hit[0], int(np.sign(arr[j]))
    assert (kind > 0).sum() == 15, kind
    # structural dependency of the 16 outputs on the 33 angles, from the reference's own autograd
    rng = np.random.RandomState(7)
    dep = np.zeros((16, 33), dtype=bool)
    for _ in range(3):
        ang = t(rng.uniform(-170, 170, (1, 33)), True)
        mm = ref.fk.Forward_Kinematics_DH_Model(rh.make_args(1), ["S1"], None)
        kw2 = dict(kw)
        kw2.update(right_leg_joints_angle=ang[:, 0:5], left_leg_joints_angle=ang[:, 5:10],
                   body_joints_angle=ang[:, 10:23], right_hand_joints_angle=ang[:, 23:28],
                   left_hand_joints_angle=ang[:, 28:33])
        w16 = mm.change_3d_joint_angle(**kw2)[:, ref.h36m.H36M_32_To_16_Table]
        for k in range(16):
            for ax in range(3):
                (g,) = torch.autograd.grad(w16[0, k, ax], ang, retain_graph=True)
                dep[k] |= (g[0].abs() > 1e-9).numpy()
    gtab = ref.generator.GAN_angle_range_table
    ranges = np.array([gtab["joint%d" % (i + 1)]["range"] for i in range(34)], dtype=np.float32)
    grtab = ref.generator.GAN_global_rotation_table
    granges = np.array([grtab["angle_" + c]["range"] for c in "xyz"], dtype=np.float32)
    subjects = ["S1", "S5", "S6", "S7", "S8", "S9", "S11"]
    cams = np.stack([np.stack([rh.camera_block(s, c) for c in range(4)]) for s in subjects])
    tmpl = np.load(os.path.join(rh.REF_ROOT, "data_extra", "bone_length_npy", "hm36s15678_bl_templates.npy"))
    return dict(alpha=alpha, theta0=theta0, len_kind=kind, len_bone=bone, len_sign=sign, dep=dep,
                h36m_32_to_16=np.array(ref.h36m.H36M_32_To_16_Table, np.int32),
                used_16key_15bone_len_table=np.array(ref.fk.used_16key_15bone_len_table, np.int32),
                gan_angle_range=ranges, gan_global_rot_range=granges, camera_blocks=cams,
                camera_subjects=np.array(subjects), bone_templates=tmpl)


def kat_fixture():
    ref = rh.import_reference()
    m = ref.fk.Forward_Kinematics_DH_Model(rh.make_args(1), ["S1"], None)
    tpose = m.init_Fk_DH_angle()
    inp = dict(ang=np.full((1, 33), 10.0, np.float32), grot=np.array([[10.0, 20.0, 30.0]], np.float32),
               bone=np.array([[.45, .45, .44, .44, .13, .13, .23, .26, .15, .15, .28, .28, .25, .25, .18]], np.float32),
               root=np.array([[1.0, 2.0, 3.0]], np.float32))
    blk = rh.camera_block("S1", 0)
    z = dict(g_world=np.ones((1, 16, 3), np.float32), g_cam=np.ones((1, 16, 3), np.float32),
             g_uv=np.ones((1, 16, 2), np.float32))
    bent = run_case(inp, blk, z)
    inp2 = dict(inp); inp2["root"] = np.array([[0.0, 0.0, 1.0]], np.float32)
    bent2 = run_case(inp2, blk, z)
    out = {"tpose32": tpose}
    out.update({"bent_" + k: v for k, v in bent.items()})
    out.update({"bent2_" + k: v for k, v in bent2.items()})
    return out


def camera_ops_fixture():
    ref = rh.import_reference()
    rng = np.random.RandomState(11)
    n = 77
    x = rng.uniform(-2, 2, (n, 16, 3)).astype(np.float32)
    x[..., 2] = rng.uniform(1.5, 6.0, (n, 16)).astype(np.float32)  # some |x/z| > 1 -> clamp active
    subjects = ["S1", "S5", "S6", "S7", "S8"]
    rows = np.stack([rh.camera_block(subjects[rng.randint(5)], rng.randint(4))[7:16] for _ in range(n)])
    rows16 = np.concatenate([rows, rng.randn(n, 7).astype(np.float32)], 1)  # 16-col variant: only 9 used
    g_uv = rng.randn(n, 16, 2).astype(np.float32)
    xt = t(x, True)
    uv = ref.camera.project_to_2d(xt, t(rows))
    (gx,) = torch.autograd.grad((uv * t(g_uv)).sum(), xt)
    uv16 = ref.camera.project_to_2d(t(x), t(rows16))
    blk = rh.camera_block("S6", 3)
    xw = t(rng.uniform(-3, 3, (n, 16, 3)), True)
    cam = ref.camera.GAN_torch_world_to_camera(xw, R=t(blk[0:4]).view(1, 4), t=t(blk[4:7]).view(1, 3))
    g_cam = rng.randn(n, 16, 3).astype(np.float32)
    (gxw,) = torch.autograd.grad((cam * t(g_cam)).sum(), xw)
    return dict(x=x, cam_rows9=rows, cam_rows16=rows16, uv=uv.detach().numpy(), uv16=uv16.detach().numpy(),
                g_uv=g_uv, g_x=gx.numpy(), clamped=(np.abs(x[..., :2] / x[..., 2:]) > 1),
                w_x=xw.detach().numpy(), w_q=blk[0:4], w_t=blk[4:7], w_cam=cam.detach().numpy(), w_g_cam=g_cam,
                w_g_x=gxw.numpy())


def sampler_fixture():
    """handler_but_generater (--data_enhancement_method normal) with a fixed seed, dataset look-ups off."""
    import argparse
    ref = rh.import_reference()
    out = {}
    for mode in ("different", "same"):
        args = argparse.Namespace(batch_size=1, random_seed=5, single_or_multi_train_mode="single", architecture="3,3,3",
                                  generator_whole_number=40, generator_choose_BoneLen=False,
                                  generator_choose_root_pos=False, generator_global_rot=True, bone_len_scaler=mode)
        m = ref.fk.Forward_Kinematics_DH_Model(args, ["S1"], None)
        m.record_bone_len = [0.45, 0.45, 0.44, 0.44, 0.13, 0.13, 0.23, 0.26, 0.15, 0.15, 0.28, 0.28, 0.25, 0.25, 0.18]
        m.root_3d_pos = np.array([0.1, -0.2, 0.9])
        pos, ang, glob, bl, root = m.handler_but_generater()
        out[mode + "_pos32"] = np.asarray(pos, np.float32)
        out[mode + "_ang"] = np.asarray(ang, np.float64)
        out[mode + "_glob"] = np.asarray(glob, np.float64)
        out[mode + "_next_draw"] = np.array([m.random.randint(0, 1 << 30)])   # RNG position after the call
    return out


def generator_fixture():
    """The reference's own Fk_Generator.forward / Video_Fk_Generator.forward (Fk_generator.py:114-261, :302-458)
    fed with a KNOWN raw last-layer output: `deconv_out` is replaced by a module that returns a leaf tensor, so
    everything after the MLP is the unmodified reference code and autograd yields d(loss)/d(raw output)."""
    import argparse
    import torch.nn as nn
    ref = rh.import_reference()

    class Feed(nn.Module):
        def __init__(self, raw):
            super().__init__()
            self.raw = raw

        def forward(self, x):
            return self.raw * 1.0

    out = {}
    rng = np.random.RandomState(21)
    from dhfk import synthetic
    for tag, B, F, pre in (("single", 70, 1, True), ("single_nopre", 33, 1, False), ("video", 4, 9, True)):
        args = argparse.Namespace(batch_size=B, random_seed=0, single_or_multi_train_mode="multi" if F > 1 else "single",
                                  architecture="3,3", GAN_OUTPUT_DIM=35, Gen_DenseDim=16, GAN_whether_use_preAngle=pre,
                                  whether_use_RT=True, bone_len_scaler="different", record_all_picture=False,
                                  checkpoint="/tmp")
        fk = ref.fk.Forward_Kinematics_DH_Model(args, ["S1"], None)
        if F == 1:
            G = ref.generator.Fk_Generator(fk, args, torch.device("cpu"))
        else:
            G = ref.generator.Video_Fk_Generator(F, fk, args, torch.device("cpu"))
        G.train_num = 1                                   # skip the heat-map dump (train_num % 500 == 1)
        raw = torch.tensor((rng.randn(B, F * 35) * 1.2).astype(np.float32), requires_grad=True)
        G.deconv_out = Feed(raw)
        bone = synthetic.gan_like(B * F, seed=31 + B)["bone"]
        G.boneLength = torch.tensor(bone)
        if F == 1:
            torch.manual_seed(77)
            fake = G(torch.zeros(B, 128))
            torch.manual_seed(77)
            scaler = (torch.randint(-200, 200, size=(B, 8)) / 1000.0).numpy()
        else:
            fk.random = np.random.RandomState(5)
            fake = G(torch.zeros(B, 128))
            scaler = np.random.RandomState(5).randint(-200, 200, size=(B, 8)).reshape(B, 8) / 1000.0
            scaler = np.repeat(scaler[:, None, :], F, axis=1).reshape(B * F, 8)
        g_fake = rng.randn(*fake.shape).astype(np.float32)
        (fake * torch.tensor(g_fake)).sum().backward()
        out.update({tag + "_raw": raw.detach().numpy(), tag + "_bone": bone, tag + "_scaler": scaler.astype(np.float32),
                    tag + "_fake": fake.detach().numpy(), tag + "_g_fake": g_fake, tag + "_d_raw": raw.grad.numpy()})
    return out


def retarget_fixture():
    """random_bl_aug (function_aug/dataloader_update.py:18-41) and video_mode_random_bl_aug
    (models_Fk_GAN/video_mode_operate.py:879-897) + project_to_2d with per-row intrinsics, run from the
    reference's own working directory (they np.load the template file by relative path)."""
    ref = rh.import_reference()
    for name in ("progress", "progress.bar"):
        if name not in sys.modules:
            rh._stub_module(name)
    cwd = os.getcwd()
    os.chdir(rh.REF_ROOT)
    try:
        from function_aug import dataloader_update as du
        # video_mode_operate drags in the whole training stack; its 12-line function is exercised through the
        # identical code path of random_bl_aug with a broadcast single row (checked below against gan_utils)
        from utils import gan_utils as gu
        cam_poses = np.load(os.path.join(OUT, "gan133.npz"))["cam"].astype(np.float32)      # realistic camera-space poses
        n = cam_poses.shape[0]
        rng = np.random.RandomState(23)
        subjects = ["S1", "S5", "S6", "S7", "S8"]
        rows = np.stack([rh.camera_block(subjects[rng.randint(5)], rng.randint(4))[7:16] for _ in range(n)])
        rows16 = np.concatenate([rows, rng.randn(n, 7).astype(np.float32)], 1)
        np.random.seed(17)
        out = du.random_bl_aug(t(cam_poses))
        np.random.seed(17)
        idx = np.random.choice(5, n)
        after = np.random.randint(0, 1 << 30)          # RNG position after the call
        uv = ref.camera.project_to_2d(out, t(rows16))
        # sequence variant: one template row for 27 frames, one shared intrinsics row (video_mode_operate.py:916-928)
        seq = cam_poses[:27]
        tm = np.load("./data_extra/bone_length_npy/hm36s15678_bl_templates.npy")
        np.random.seed(4)
        vidx = np.random.choice(tm.shape[0], 1)
        x = t(seq)
        root = x[:, :1, :] * 1.0
        unit = gu.get_bone_unit_vecbypose3d(x - x[:, :1, :])
        vout = gu.get_pose3dbyBoneVec(unit * torch.from_numpy(tm[vidx].astype("float32")).unsqueeze(2)) + root
        used = torch.zeros(27, 9)
        used[:] = t(rows[0])
        vuv = ref.camera.project_to_2d(vout, used)
    finally:
        os.chdir(cwd)
    return dict(pose=cam_poses, tmpl_idx=idx.astype(np.int32), rng_after=np.array([after]), cam_rows16=rows16,
                out_pose=out.numpy(), out_uv=uv.numpy(), templates=tm, v_pose=seq, v_idx=vidx.astype(np.int32),
                v_cam_row=rows[0], v_out_pose=vout.numpy(), v_out_uv=vuv.numpy())


def critic_fixture():
    import argparse
    rh.import_reference()
    for name in ("progress", "progress.bar"):
        if name not in sys.modules:
            rh._stub_module(name)
    from models_Fk_GAN import Fk_discriminator as fd
    cpu = torch.device("cpu")
    g = np.load(os.path.join(OUT, "gan133.npz"))
    world = g["world16"].astype(np.float32)
    n = world.shape[0]
    rng = np.random.RandomState(41)
    g_pos = rng.randn(n, 16, 3).astype(np.float32)
    g_kcs = rng.randn(n, 30).astype(np.float32)
    out = dict(pose=world, g_pos=g_pos, g_kcs=g_kcs)
    left, right = [4, 5, 6, 10, 11, 12], [1, 2, 3, 13, 14, 15]

    def transform(x, centre, flip):
        if centre:
            x = x[:, :, :] - x[:, :1, :]                      # model_fk_gan_train.py:312
        if flip:                                              # :325-327 (on a clone; autograd-friendly restatement
            x = x * torch.tensor([-1.0, 1.0, 1.0])            #  of `[:, :, 0] *= -1`)
            idx = list(range(16))
            for d, s_ in zip(left + right, right + left):
                idx[d] = s_
            x = x[:, idx, :]
        return x

    for centre in (0, 1):
        for flip in (0, 1):
            tag = "c%df%d" % (centre, flip)
            x = t(world, True)
            pos = transform(x, centre, flip)
            kcs = fd.special_KCS_Input_transform(pos, cpu)
            vk = fd.video_mode_special_KCS_Input_transform(pos, cpu)
            out[tag + "_pos"] = pos.detach().numpy()
            out[tag + "_kcs"] = kcs.detach().numpy()
            out[tag + "_vkcs"] = vk.detach().numpy()
            (gx,) = torch.autograd.grad((pos * t(g_pos)).sum() + (kcs * t(g_kcs)).sum(), x, retain_graph=True)
            out[tag + "_g_pose"] = gx.numpy()
            (gk,) = torch.autograd.grad((kcs * t(g_kcs)).sum(), x, retain_graph=True)
            out[tag + "_g_pose_kcs_only"] = gk.numpy()
            (gv,) = torch.autograd.grad((vk * t(g_kcs[:, :15])).sum(), x)
            out[tag + "_g_pose_vkcs_only"] = gv.numpy()
    # the in-place flip exactly as the train loop writes it, 3-D and 2-D (model_fk_gan_train.py:324-327, 399-405)
    f3 = t(world).detach().clone()
    f3[:, :, 0] *= -1
    f3[:, left + right, :] = f3[:, right + left, :]
    uv = g["uv"].astype(np.float32)
    f2 = t(uv).detach().clone()
    f2[:, :, 0] *= -1
    f2[:, left + right, :] = f2[:, right + left, :]
    out.update(flip3=f3.numpy(), uv=uv, flip2=f2.numpy())
    # reference 3-D critic, seeded weights: outputs, WGAN-GP penalty and its parameter gradients
    B = 40
    args = argparse.Namespace(Dis_DenseDim_3D=32)
    torch.manual_seed(12)
    D = fd.Fk_3D_Discriminator(cpu, args)
    real = t(world[:B]) - t(world[:B])[:, :1]
    fake = t(world[B:2 * B]) - t(world[B:2 * B])[:, :1]
    d_real = D(real)
    D.zero_grad()
    torch.manual_seed(3)
    gp = fd.calc_gradient_penalty(D, real.data, fake.data, B, 10, cpu)
    gp.backward()
    torch.manual_seed(3)
    alpha = torch.rand(B, 1)
    names = [k for k, _ in D.named_parameters()]
    out.update(d3d_real=real.numpy(), d3d_fake=fake.numpy(), d3d_out_real=d_real.detach().numpy(),
               d3d_gp=np.array([gp.item()], np.float64), d3d_alpha=alpha.numpy(),
               d3d_param_names=np.array(names))
    for k, v in D.state_dict().items():
        out["d3d_w_" + k] = v.numpy()
    for k, v in D.named_parameters():
        out["d3d_g_" + k] = v.grad.numpy() if v.grad is not None else np.zeros(tuple(v.shape), np.float32)
    return out


def video_critic_fixture():
    """Fk_discriminator.py:381-587 and video_GAN_fun.py:222-223,269-270 on B=12 clips of F=9 frames."""
    import argparse
    rh.import_reference()
    for name in ("progress", "progress.bar"):
        if name not in sys.modules:
            rh._stub_module(name)
    from models_Fk_GAN import Fk_discriminator as fd
    cpu = torch.device("cpu")
    F, B = 9, 12
    g = np.load(os.path.join(OUT, "gan133.npz"))
    world = g["world16"].astype(np.float32)[:B * F]
    world = world - world[:, :1]                                     # root-centred, as the loop feeds them (:206)
    # make the clips temporally coherent enough to be meaningful: frame f = blend towards the clip's first pose
    clips = world.reshape(B, F, 16, 3).copy()
    w = np.linspace(0.0, 0.6, F, dtype=np.float32).reshape(1, F, 1, 1)
    clips = (1 - w) * clips + w * clips[:, :1]
    uv = g["uv"].astype(np.float32)[:B * F].reshape(B, F, 16, 2)
    uv = (1 - w) * uv + w * uv[:, :1]
    out = dict(pose=clips.reshape(B * F, 16, 3), uv=uv.reshape(B * F, 16, 2), frames=np.array([F]))
    args = argparse.Namespace(video_Dis_DenseDim_3D=8, video_Dis_DenseDim_2D=8,
                              motion_Dis_whether_use_3dPos_branch=True, motion_Dis_whether_use_3dDiff_branch=True)

    def run(tag, D, x_np, width, branch_inputs, real_np, fake_np):
        feats = {}
        hooks = [getattr(D, name).register_forward_pre_hook(
            lambda m, inp, key=key: feats.__setitem__(key, inp[0].detach().numpy().copy()))
            for key, name in branch_inputs.items()]
        x = t(x_np, True)
        d = D(x)
        for h in hooks:
            h.remove()
        for key, v in feats.items():
            if key != "pos":                                         # the position branches receive x itself
                out["%s_feat_%s" % (tag, key)] = v
            else:
                assert np.array_equal(v.reshape(x_np.shape), x_np)
        rng = np.random.RandomState(7)
        gd = rng.randn(*d.shape).astype(np.float32)
        (gx,) = torch.autograd.grad((d * t(gd)).sum(), x)
        out[tag + "_out"] = d.detach().numpy()
        out[tag + "_g_out"] = gd
        out[tag + "_g_x"] = gx.numpy()
        D.zero_grad()
        torch.manual_seed(3)
        gp = fd.calc_gradient_penalty(D, t(real_np).data, t(fake_np).data, real_np.shape[0], 10, cpu)
        gp.backward()
        out[tag + "_gp"] = np.array([gp.item()], np.float64)
        for k, v in D.named_parameters():
            out["%s_gpgrad_%s" % (tag, k)] = v.grad.numpy().copy() if v.grad is not None else \
                np.zeros(tuple(v.shape), np.float32)

    torch.manual_seed(21)
    D3 = fd.Video_motion_Fk_3D_Discriminator(cpu, args, F)
    for k, v in D3.state_dict().items():
        out["d3_w_" + k] = v.numpy().copy()
    x3 = clips.reshape(B, F, 48)
    half = B // 2
    b3 = dict(kcs="special_KCS_previous", dkcs="diff_special_KCS_previous", pos="pos_3d_previous",
              dpos="diff_pos_3d_previous")
    run("d3_fwd", D3, x3, 48, b3, x3[:half], x3[half:])
    x3r = torch.clone(torch.flip(t(x3), dims=[1])).numpy()               # video_GAN_fun.py:222-223
    run("d3_rev", D3, x3r, 48, b3, x3r[:half], x3r[half:])
    torch.manual_seed(22)
    D2 = fd.Video_motion_Fk_2D_Discriminator(cpu, args, F)
    for k, v in D2.state_dict().items():
        out["d2_w_" + k] = v.numpy().copy()
    x2 = uv.reshape(B, F, 32)
    b2 = dict(pos="pos_2d_previous", rdiff="root_diff_2d_previous")
    run("d2_fwd", D2, x2, 32, b2, x2[:half], x2[half:])
    x2r = torch.clone(torch.flip(t(x2), dims=[1])).numpy()               # video_GAN_fun.py:269-270
    run("d2_rev", D2, x2r, 32, b2, x2r[:half], x2r[half:])
    torch.manual_seed(3)
    out["gp_alpha"] = torch.rand(half, 1).numpy()
    return out


def gan_loop_fixture():
    """model_fk_gan_train.py:236-512 and video_GAN_fun.py:79-602, unmodified, with torch.device("cuda") -> CPU."""
    import ref_loop
    out = {}
    for mode, batch in (("single", 32), ("video", 8)):
        r = ref_loop.run_loop(mode, device="cpu", iters=6, batch=batch, dense=16, seed=11)
        names, vals = ref_loop.scalars_matrix(r["scalars"])
        out[mode + "_scalar_names"] = np.array(names)
        out[mode + "_scalars"] = vals
        out[mode + "_g_grads"] = np.stack(r["g_grads"]).astype(np.float32)
        out[mode + "_buffer_3d"] = np.asarray(r["buffer_3d"], np.float32)
        out[mode + "_buffer_2d"] = np.asarray(r["buffer_2d"], np.float32)
        out[mode + "_model_G"] = r["params"]["model_G"].astype(np.float32)
        out[mode + "_cfg"] = np.array([6, batch, 16, 11])
    return out


def fixtures_table():
    from dhfk import synthetic
    return {
        "tables": tables_fixture,
        "kat": kat_fixture,
        "gan133": lambda: run_case(synthetic.gan_like(133, seed=1234), rh.camera_block("S1", 0),
                                   synthetic.upstream_grads(133)),
        "stress200": lambda: run_case(synthetic.gan_like(200, seed=99, root_mode="generator", angle_mode="stress"),
                                      rh.camera_block("S7", 2), synthetic.upstream_grads(200, seed=5)),
        "video36": lambda: run_case(synthetic.gan_like(36, seed=3), rh.camera_block("S8", 1),
                                    synthetic.upstream_grads(36, seed=8), mode="multi", architecture="3,3",
                                    root_shape=(4, 9, 3)),
        "camera_ops": camera_ops_fixture,
        "sampler40": sampler_fixture,
        "generator": generator_fixture,
        "retarget": retarget_fixture,
        "critic": critic_fixture,
        "video_critic": video_critic_fixture,
        "gan_loop": gan_loop_fixture,
    }


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    fixtures = fixtures_table()
    for name in (sys.argv[1:] or list(fixtures)):
        np.savez(os.path.join(OUT, name + ".npz"), **fixtures[name]())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
