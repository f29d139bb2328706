"""Drop-in for the critic side of the reference's ``models_Fk_GAN/Fk_discriminator.py``.

Input transforms -- these are the kernels:
  ``special_KCS_Input_transform`` (Fk_discriminator.py:36-146) and ``video_mode_special_KCS_Input_transform``
  (:269-377) keep their names and argument meaning and run as one sm_100a kernel each way (forward, vector-Jacobian
  backward, and the Jacobian-vector product that WGAN-GP's ``create_graph=True`` pass differentiates through).
  ``video_motion_3d_forward`` / ``video_motion_2d_forward`` are the ``forward`` methods of the two motion critics
  (:436-512, :554-587) with the per-frame KCS, the F-1 slice-write loops of the adjacent-frame differences and the
  clones replaced by ONE launch (``functional.video_critic_input`` / ``video_root_diff``).

Critic networks -- plain Linear / ReLU stacks, cuBLAS-backed torch as the north star says.  The reference's classes
look the transforms up as module globals, so ``dropin.install(critics=True)`` rebinds the two functions and the two
``forward`` methods on the reference's own classes and nothing else.  The classes below are the same networks
(constructor arguments, sub-module names and state dicts are interchangeable with the reference's) for use without
the reference tree: tools/gan_step_bench.py, bench.py's GAN-step extras and the GPU tests.

``critic_views`` is the opt-in fused form of the train loop's root-centring + flip (model_fk_gan_train.py:311-331).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .Fk_generator import myResNet
from .functional import critic_input, flip_pose, video_critic_input, video_root_diff  # noqa: F401


def special_KCS_Input_transform(pos_16_3d, device=None):
    """pos_16_3d [N,16,3] or [N,48] -> [N,30]: 15 bone-pair cosines then 15 bone lengths."""
    return critic_input(pos_16_3d.view(-1, 16, 3), kcs_cols=30, return_pos=False)


def video_mode_special_KCS_Input_transform(pos_16_3d, device=None):
    """pos_16_3d [N,16,3] or [N,48] -> [N,15]: the 15 bone-pair cosines."""
    return critic_input(pos_16_3d.view(-1, 16, 3), kcs_cols=15, return_pos=False)


def critic_views(pose16, flip=True):
    """(centred, flipped-and-centred or None) views of a [N,16,3] batch as the train loop builds them for the 3-D
    critic: `x - x[:, :1]`, then on a detached clone `[:, :, 0] *= -1` and the left/right joint swap."""
    centred = critic_input(pose16, centre=True, kcs_cols=0)
    flipped = critic_input(pose16.detach(), centre=True, flip=True, kcs_cols=0) if flip else None
    return centred, flipped


def _stack(x, first, *blocks):
    x = first(x)
    for b in blocks:
        x = b(x)
    return x


def video_motion_3d_forward(self, input, reverse=False):
    """Video_motion_Fk_3D_Discriminator.forward (Fk_discriminator.py:436-512).  `self` needs the reference's
    sub-modules and `video_frame_num` / `args`.  reverse=True evaluates the critic on the clip played backwards
    (what the train loop feeds after torch.flip(x, dims=[1]), video_GAN_fun.py:222-223) without flipping it."""
    F = self.video_frame_num
    use_pos = bool(self.args.motion_Dis_whether_use_3dPos_branch)
    use_diff = bool(self.args.motion_Dis_whether_use_3dDiff_branch)
    x = input.reshape(-1, 16, 3)
    outs = video_critic_input(x, F, reverse=reverse, want_dpos=use_diff and F > 1, want_pos=use_pos and reverse)
    kcs, dkcs = outs[0], outs[1]
    rest = list(outs[2:])
    feats = [_stack(kcs.view(-1, F * 15), self.special_KCS_previous, self.special_KCS_block1, self.special_KCS_block2,
                    self.special_KCS_block3),
             _stack(dkcs.view(-1, (F - 1) * 15), self.diff_special_KCS_previous, self.diff_special_KCS_block1,
                    self.diff_special_KCS_block2, self.diff_special_KCS_block3)]
    dpos = rest.pop(0) if (use_diff and F > 1) else None
    if use_pos:
        pos = rest.pop(0) if reverse else input
        feats.append(_stack(pos.reshape(-1, F * 48), self.pos_3d_previous, self.pos_3d_block1, self.pos_3d_block2,
                            self.pos_3d_block3))
    if use_diff:
        if dpos is None:
            dpos = x.new_zeros((x.shape[0] // F, 0))
        feats.append(_stack(dpos.view(-1, (F - 1) * 48), self.diff_pos_3d_previous, self.diff_pos_3d_block1,
                            self.diff_pos_3d_block2, self.diff_pos_3d_block3))
    out = torch.cat(feats, dim=-1)
    return self.kcs_output(self.kcs_merge_block1(self.kcs_merge_previous(out)))


def video_motion_2d_forward(self, input, reverse=False):
    """Video_motion_Fk_2D_Discriminator.forward (Fk_discriminator.py:554-587)."""
    F = self.video_frame_num
    if reverse:
        diff, pos = video_root_diff(input.reshape(-1, 16, 2), F, reverse=True, want_playback=True)
    else:
        diff, pos = video_root_diff(input.reshape(-1, 16, 2), F), input
    a = _stack(pos.reshape(-1, F * 32), self.pos_2d_previous, self.pos_2d_block1, self.pos_2d_block2, self.pos_2d_block3)
    b = _stack(diff.view(-1, (F - 1) * 2), self.root_diff_2d_previous, self.root_diff_2d_block1, self.root_diff_2d_block2,
               self.root_diff_2d_block3)
    return self.merge_output(self.merge_block1(self.merge_previous(torch.cat((a, b), dim=-1))))


def _branch(owner, prefix, in_dim, dim, first="previous"):
    """`<prefix><first>` = Linear+ReLU, `<prefix>block1..3` = residual blocks: the reference's per-branch layout."""
    setattr(owner, prefix + first, nn.Sequential(nn.Linear(in_dim, dim), nn.ReLU(True)))
    for i in (1, 2, 3):
        setattr(owner, "%sblock%d" % (prefix, i), myResNet(dim))


class Fk_3D_Discriminator(nn.Module):
    """Single-frame 3-D critic (Fk_discriminator.py:149-203): position branch + KCS-30 branch."""

    def __init__(self, device, args):
        super().__init__()
        self.device, self.args = device, args
        dim = args.Dis_DenseDim_3D
        _branch(self, "", 48, dim)
        _branch(self, "special_KCS_", 30, dim)
        self.merge_previous = nn.Sequential(nn.Linear(2 * dim, 100), nn.ReLU(True))
        self.merge_block1 = myResNet(100)
        self.output = nn.Linear(100, 1)

    def forward(self, input):
        kcs = special_KCS_Input_transform(input, self.device)
        a = _stack(kcs, self.special_KCS_previous, self.special_KCS_block1, self.special_KCS_block2, self.special_KCS_block3)
        b = _stack(input.reshape(-1, 48), self.previous, self.block1, self.block2, self.block3)
        return self.output(self.merge_block1(self.merge_previous(torch.cat((a, b), dim=-1))))


class Fk_2D_Discriminator(nn.Module):
    """Single-frame 2-D critic (Fk_discriminator.py:239-266)."""

    def __init__(self, args, num_joints=16):
        super().__init__()
        self.args = args
        dim = args.Dis_DenseDim_2D
        self.pose_layer_1 = nn.Linear(num_joints * 2, dim)
        self.pose_layer_2 = nn.Linear(dim, dim)
        self.pose_layer_3 = nn.Linear(dim, dim)
        self.pose_layer_4 = nn.Linear(dim, dim)
        self.layer_last = nn.Linear(dim, dim)
        self.layer_pred = nn.Linear(dim, 1)
        self.relu = nn.LeakyReLU()

    def forward(self, x):
        x = x.reshape(-1, 32)
        d1 = self.relu(self.pose_layer_1(x))
        d2 = self.relu(self.pose_layer_2(d1))
        d3 = self.relu(self.pose_layer_3(d2) + d1)
        d4 = self.pose_layer_4(d3)
        return self.layer_pred(self.relu(self.layer_last(d4)))


class Video_motion_Fk_3D_Discriminator(nn.Module):
    """Motion 3-D critic (Fk_discriminator.py:381-512)."""

    def __init__(self, device, args, video_frame_num):
        super().__init__()
        self.video_frame_num, self.device, self.args = video_frame_num, device, args
        F, dim = video_frame_num, args.video_Dis_DenseDim_3D
        _branch(self, "special_KCS_", F * 15, dim)
        _branch(self, "diff_special_KCS_", (F - 1) * 15, dim)
        _branch(self, "pos_3d_", F * 48, dim)
        _branch(self, "diff_pos_3d_", (F - 1) * 48, dim)
        self.branch_num = 2 + int(bool(args.motion_Dis_whether_use_3dPos_branch)) + \
            int(bool(args.motion_Dis_whether_use_3dDiff_branch))
        self.kcs_merge_previous = nn.Sequential(nn.Linear(dim * self.branch_num, 100), nn.ReLU(True))
        self.kcs_merge_block1 = myResNet(100)
        self.kcs_output = nn.Linear(100, 1)

    forward = video_motion_3d_forward


class Video_motion_Fk_2D_Discriminator(nn.Module):
    """Motion 2-D critic (Fk_discriminator.py:516-587)."""

    def __init__(self, device, args, video_frame_num):
        super().__init__()
        self.video_frame_num, self.device, self.args = video_frame_num, device, args
        F, dim = video_frame_num, args.video_Dis_DenseDim_2D
        _branch(self, "pos_2d_", F * 32, dim)
        _branch(self, "root_diff_2d_", (F - 1) * 2, dim)
        self.merge_previous = nn.Sequential(nn.Linear(2 * dim, 100), nn.ReLU(True))
        self.merge_block1 = myResNet(100)
        self.merge_output = nn.Linear(100, 1)

    forward = video_motion_2d_forward
