"""Drop-in for the per-epoch loader refresh of the reference.

Reference-shaped names:
  random_bl_aug(x)                      function_aug/dataloader_update.py:18-41
  video_mode_random_bl_aug(x)           models_Fk_GAN/video_mode_operate.py:879-897
  dataloader_update(args, data_dict, device)          function_aug/dataloader_update.py:43-107
  video_mode_dataloader_update(args, data_dict, device)   models_Fk_GAN/video_mode_operate.py:898-968
  refresh_poses(poses, cams, ...)       the same work for a device-resident pose bank (no loader round trip)

The reference root-centres each pose, takes unit bone vectors with two [N,3,16]x[N,16,15] matmuls, multiplies
by a random bone-length template row, rebuilds the pose with a third matmul, then projects with per-row
intrinsics (~25 torch ops and one D2H copy per batch).  Here one kernel (dhfk_retarget_project) does all of
it per pose in registers; the host side keeps the reference's random stream (np.random.choice per batch) so a
seeded run picks the same template rows.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import tables
from .functional import retarget_project

_TEMPLATE_FILE = "./data_extra/bone_length_npy/hm36s15678_bl_templates.npy"


def bone_length_templates() -> np.ndarray:
    """(T,15) float32 in utils/gan_utils.py bone order.  The reference reads the .npy relative to the working
    directory on every call (dataloader_update.py:24); when that file is present it wins, otherwise the copy of
    its 5x15 values kept in tables.py is used."""
    if os.path.exists(_TEMPLATE_FILE):
        return np.load(_TEMPLATE_FILE).astype("float32")
    return tables.BONE_TEMPLATES_GANUTILS_ORDER


def random_bl_aug(x):
    """x [N,16,3] -> [N,16,3]: every pose gets the bone lengths of a random S1/5/6/7/8 template row, joint
    directions and root position kept.  Consumes np.random exactly like the reference (one choice(T, N))."""
    tm = bone_length_templates()
    tmp_idx = np.random.choice(tm.shape[0], x.shape[0])
    return retarget_project(x, tm, tmp_idx)


def video_mode_random_bl_aug(x):
    """Sequence variant: ONE template row for all frames of the sequence x [F,16,3] (choice(T, 1))."""
    tm = bone_length_templates()
    tmp_idx = np.random.choice(tm.shape[0], 1)
    return retarget_project(x, tm[tmp_idx])


def refresh_poses(poses16, cam_rows, templates=None, tmpl_idx=None, out_pose=None, out_uv=None):
    """Device-resident variant: poses16 [N,16,3] and cam_rows [N,9|16] already on the GPU; returns
    (poses', uv') on the GPU, optionally into caller-provided buffers (a fake/real pose bank that is never
    copied back to the host)."""
    tm = bone_length_templates() if templates is None else templates
    if tmpl_idx is None:
        tmpl_idx = np.random.choice(np.asarray(tm).shape[0] if not isinstance(tm, torch.Tensor) else tm.shape[0],
                                    poses16.shape[0])
    return retarget_project(poses16, tm, tmpl_idx, cam_rows, out_pose=out_pose, out_uv=out_uv)


def dataloader_update(args, data_dict, device):
    """Same contract as the reference's dataloader_update: rebuilds data_dict['train_gt2d3d_loader'],
    ['target_3d_loader'] and ['target_2d_loader'] from bone-length-swapped poses and their re-projections.
    One fused launch per batch; results leave the GPU once, at the end, instead of three .cpu() calls per batch.
    The Dataset / DataLoader classes are the reference's own (imported at call time)."""
    from torch.utils.data import DataLoader
    from common.data_loader import PoseDataSet, PoseTarget  # reference module (the caller's sys.path)

    tm = bone_length_templates()
    poses, uvs, cams, actions = [], [], [], []
    for targets_3d, _, action, cam_param in data_dict["train_gt2d3d_loader"]:
        targets_3d = targets_3d.to(device, non_blocking=True)
        cam_param = cam_param.to(device, non_blocking=True)
        tmp_idx = np.random.choice(tm.shape[0], targets_3d.shape[0])
        p, uv = retarget_project(targets_3d, tm, tmp_idx, cam_param)
        poses.append(p)
        uvs.append(uv)
        cams.append(cam_param)
        actions.append(action)
    assert len(poses) == len(uvs) == len(actions) == len(cams)
    sizes = [p.shape[0] for p in poses]
    split = lambda ts: list(np.split(torch.cat(ts).cpu().numpy(), np.cumsum(sizes)[:-1])) if ts else []
    buffer_poses_train, buffer_poses_train_2d, buffer_cams_train = split(poses), split(uvs), split(cams)
    print("==> Random Bone Length (S15678) swap completed")
    kw = dict(batch_size=args.batch_size, shuffle=True, num_workers=args.num_workers, pin_memory=True)
    data_dict["train_gt2d3d_loader"] = DataLoader(
        PoseDataSet(buffer_poses_train, buffer_poses_train_2d, actions, buffer_cams_train), **kw)
    data_dict["target_3d_loader"] = DataLoader(PoseTarget(buffer_poses_train), **kw)
    data_dict["target_2d_loader"] = DataLoader(PoseTarget(buffer_poses_train_2d), **kw)
    return


def video_mode_dataloader_update(args, data_dict, device):
    """Same contract as the reference's video_mode_dataloader_update: every training sequence gets the bone lengths of
    ONE random template row (choice(T, 1) per sequence, in sequence order -- the reference's random stream) and is
    re-projected with its own camera; data_dict['target_GAN_loader'] is rebuilt from the result.  The reference makes
    ~25 torch calls and three .cpu() copies PER SEQUENCE; here the whole training set goes through ONE fused launch
    (per-pose template index, per-pose intrinsics row) and comes back once.  The chunked generator class is the
    reference's own (imported at call time)."""
    import copy

    from models_Fk_GAN.video_mode_operate import GAN_video_ChunkedGenerator, video_receptive_field  # reference module

    tm = bone_length_templates()
    poses = [np.asarray(p, dtype="float32") for p in data_dict["poses_train"]]
    cams = [np.asarray(c, dtype="float32") for c in data_dict["cams_train"]]
    actions = list(data_dict["actions_train"])
    assert len(poses) == len(cams) == len(actions)
    sizes = [p.shape[0] for p in poses]
    rows = np.concatenate([np.full(k, np.random.choice(tm.shape[0], 1)[0], dtype=np.int32) for k in sizes])
    cam_rows = np.concatenate([np.repeat(c[None, :9], k, axis=0) for c, k in zip(cams, sizes)])
    out_pose, out_uv = retarget_project(torch.from_numpy(np.concatenate(poses)).to(device), tm, rows,
                                        torch.from_numpy(cam_rows).to(device))
    cuts = np.cumsum(sizes)[:-1]
    buffer_poses_train = list(np.split(out_pose.cpu().numpy(), cuts))
    buffer_poses_train_2d = list(np.split(out_uv.cpu().numpy(), cuts))
    buffer_cams_train = cams
    joints_left = out_left = [4, 5, 6, 10, 11, 12]
    joints_right = out_right = [1, 2, 3, 13, 14, 15]
    pad = (video_receptive_field([int(x) for x in args.architecture.split(",")]) - 1) // 2
    data_dict["target_GAN_loader"] = GAN_video_ChunkedGenerator(
        args.batch_size // 1, copy.deepcopy(buffer_cams_train), copy.deepcopy(buffer_poses_train),
        copy.deepcopy(buffer_poses_train_2d), chunk_length=1, pad=pad, causal_shift=0, shuffle=True, augment=False,
        kps_left=out_left, kps_right=out_right, joints_left=joints_left, joints_right=joints_right)
    return
