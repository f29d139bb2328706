"""Drop-in for ``models_Fk_GAN/Fk_generator.py``: same classes, constructor arguments, sub-module names
(state dicts are interchangeable) and return shapes; the MLP stays on cuBLAS-backed torch and everything
after the last Linear layer (Fk_generator.py:121-259 / :310-456: two tanh, 37 + 37 in-place column writes, the
range map, 15 length products, FK, the 32->16 gather -- ~400 autograd ops in the reference) is ONE fused
kernel launch (`dhfk_generator_forward`), plus one elementwise op for the bone-length scaler.

Reference quirks handled on purpose: `bone_len_scaler` 'same' / '' crash in the reference's
single-frame generator when CUDA is available (numpy `.to`); here they work as the video generator defines
them.  The every-500-calls heat-map dump (:173-177) is not reproduced.  `whether_use_RT=False` zeroes the
global rotation like the reference (:187-189)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import tables
from .functional import generator_fk

GAN_global_rotation_table = {"angle_" + c: {"range": (-180, 180), "changeRate": (-5, 5)} for c in "xyz"}
GAN_angle_range_table = {
    "joint%d" % (i + 1): {"range": (int(lo), int(hi))} for i, (lo, hi) in enumerate(tables.GAN_ANGLE_RANGE)}

_GROUP = torch.as_tensor(tables.BONE_SCALER_GROUP)
_GROUP_ON = {}      # per-device copies (a host->device copy per call would also break CUDA-graph capture)


def _group_on(device):
    g = _GROUP_ON.get(device)
    if g is None:
        idx = _GROUP.to(device)
        g = _GROUP_ON[device] = (idx.clamp(min=0), (idx >= 0).to(torch.float32))
    return g


class myResNet(nn.Module):
    """models_Fk_GAN/special_operate.py:490-510"""

    def __init__(self, DIM):
        super().__init__()
        self.fc1 = nn.Linear(DIM, DIM)
        self.fc2 = nn.Linear(DIM, DIM)
        self.relu = nn.ReLU(True)

    def forward(self, input):
        output = self.relu(self.fc1(input))
        output = self.fc2(output)
        output = output + input
        return self.relu(output)


def bone_vectors_to_lengths(pose16):
    """15 bone lengths in used_16key_15bone_len_table order from [N,16,3] poses
    (Fk_get_boneVecByPose3d + norm, special_operate.py:513-539, Fk_generator.py:107-111)."""
    idx = torch.as_tensor(tables.used_16key_15bone_len_table, device=pose16.device)
    return (pose16[:, idx[:, 1]] - pose16[:, idx[:, 0]]).norm(dim=-1)


def scaled_bone_lengths(bone, scaler):
    """boneLength[:, i] * (1 + scaler[:, group(i)]), thorax unscaled (Fk_generator.py:216-230)."""
    col, scaled = _group_on(bone.device)
    return bone * (1.0 + scaler[:, col] * scaled)


class _GeneratorBase(nn.Module):
    def _init_common(self, FK_DH_Class, args, device, INPUT_VEC_DIM, out_dim):
        self.OUTPUT_DIM = args.GAN_OUTPUT_DIM
        if self.OUTPUT_DIM != 35:
            raise ValueError("GAN_OUTPUT_DIM must be 35 (32 + 3), as in function_aug/config.py:85")
        self.BATCH_SIZE = args.batch_size
        self.FK_DH_Class = FK_DH_Class
        self.train_num = 0
        self.args = args
        self.INPUT_VEC_DIM = INPUT_VEC_DIM
        self.boneLength = torch.zeros((self.BATCH_SIZE, 15), dtype=torch.float32)
        self.device = device
        self.distribute_angle = []
        dense = args.Gen_DenseDim
        self.preprocess = nn.Sequential(nn.Linear(INPUT_VEC_DIM, dense), nn.ReLU(True))
        self.block1 = myResNet(dense)
        self.block2 = myResNet(dense)
        self.block3 = myResNet(dense)
        self.deconv_out = nn.Linear(dense, out_dim)
        self.sigmoid = nn.Sigmoid()
        self.Tanh = nn.Tanh()

    def _mlp(self, input):
        output = self.preprocess(input)
        output = self.block1(output)
        output = self.block2(output)
        output = self.block3(output)
        return self.deconv_out(output)

    def _slot_scale(self):
        half, mid = tables.generator_slot_scale(bool(getattr(self.args, "GAN_whether_use_preAngle", True)))
        if not getattr(self.args, "whether_use_RT", True):
            half, mid = half.copy(), mid.copy()
            half[34:] = 0.0   # global rotation forced to zero (Fk_generator.py:187-189)
            mid[34:] = 0.0
        return half, mid

    # Optional hook: a callable (rows, frames) -> [rows*frames, 8] tensor already on the compute device.  The reference
    # draws the scalers on the HOST every call (below); a caller that captures the training step in a CUDA graph sets
    # this to a device-side draw (no host RNG, no H2D copy inside the captured region).
    scaler_source = None

    def _draw_scaler(self, rows, frames):
        """[rows*frames, 8] scaler, RNG use as in the reference (torch.randint on the global CPU generator for the
        single-frame 'different' mode, :197; FK_DH_Class.random otherwise, :201,:383-393)."""
        mode = self.args.bone_len_scaler
        if mode == "different":
            if frames == 1:
                s = torch.randint(-200, 200, size=(rows, 8)) / 1000.0
            else:
                s = torch.tensor(np.array(self.FK_DH_Class.random.randint(-200, 200, size=(rows, 8))).reshape(rows, 8)
                                 / 1000.0, dtype=torch.float32)
                s = s.unsqueeze(1).repeat(1, frames, 1).reshape(rows * frames, 8)
        elif mode == "same":
            s = np.array(self.FK_DH_Class.random.randint(-200, 200, size=(rows, 8))).reshape(rows, 8)[:, :1]
            s = torch.tensor(np.repeat(s, 8, axis=1) / 1000.0, dtype=torch.float32)
            s = s.unsqueeze(1).repeat(1, frames, 1).reshape(rows * frames, 8)
        elif mode == "":
            s = torch.zeros((rows * frames, 8))
        else:
            raise ValueError("args.bone_len_scaler")
        return s


class Fk_Generator(_GeneratorBase):
    """Single-frame generator, models_Fk_GAN/Fk_generator.py:79-261."""

    def __init__(self, FK_DH_Class, args, device, INPUT_VEC_DIM=128):
        super().__init__()
        self._init_common(FK_DH_Class, args, device, INPUT_VEC_DIM, args.GAN_OUTPUT_DIM)

    def GAN_generator_get_bone_length(self, input):
        self.boneLength = bone_vectors_to_lengths(input.view(-1, 16, 3))

    def forward(self, input):
        net_out = self._mlp(input)                                   # [B, 35] raw
        self.train_num += 1
        scaler = (self.scaler_source(net_out.shape[0], 1) if self.scaler_source is not None
                  else self._draw_scaler(net_out.shape[0], 1).to(net_out.device))
        bone = scaled_bone_lengths(self.boneLength.to(net_out.device), scaler)
        half, mid = self._slot_scale()
        world16 = generator_fk(net_out, bone, half37=half, mid37=mid, root_scale=10.0)
        return world16.reshape(-1, 16 * 3)


class Video_Fk_Generator(_GeneratorBase):
    """Multi-frame generator, models_Fk_GAN/Fk_generator.py:264-458: the network emits F*35 columns per clip,
    frames are folded into the pose batch (N = B*F), scalers are drawn per clip and repeated over frames."""

    def __init__(self, video_frame_num, FK_DH_Class, args, device, INPUT_VEC_DIM=128):
        super().__init__()
        self.video_frame_num = video_frame_num
        self._init_common(FK_DH_Class, args, device, INPUT_VEC_DIM, video_frame_num * args.GAN_OUTPUT_DIM)

    def GAN_generator_get_bone_length(self, input):
        self.boneLength = bone_vectors_to_lengths(input.view(-1, 16, 3)).view(-1, 15)

    def forward(self, input):
        net_out = self._mlp(input).contiguous().view(-1, self.OUTPUT_DIM)   # [B*F, 35] raw
        self.train_num += 1
        rows = net_out.shape[0] // self.video_frame_num
        scaler = (self.scaler_source(rows, self.video_frame_num) if self.scaler_source is not None
                  else self._draw_scaler(rows, self.video_frame_num).to(net_out.device))
        bone = scaled_bone_lengths(self.boneLength.to(net_out.device), scaler)
        half, mid = self._slot_scale()
        world16 = generator_fk(net_out, bone, half37=half, mid37=mid, root_scale=10.0)
        fake = world16.reshape(-1, 16 * 3)
        if self.video_frame_num > 1:
            fake = fake.view(rows, self.video_frame_num, 16 * 3)
        return fake
