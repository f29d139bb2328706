"""Drop-in for ``models_Fk_GAN/forward_kinematics_DH_model.py`` (the hot part).

``Forward_Kinematics_DH_Model`` keeps the reference's constructor, attributes and the 22-kwarg
``change_3d_joint_angle`` (forward_kinematics_DH_model.py:354-364), but the torch branch
(:562-822: 33 dh_matrix builds, 46 bmm, 51 column scatters, ~6.9k ATen calls) is one fused
sm_100a kernel launch plus a 32-slot scatter, and its backward is one more launch.

Differences from the reference, all deliberate:
  * N is taken from the inputs, not from ``args.batch_size * F`` baked in at construction.
  * No O(N) Python table replication at construction, no persistent [N,n] tables, no in-place
    bone-length writes -> bone lengths may require grad and calls are stateless.
  * Non-tensor inputs (the reference's numpy branch, :366-560, used by ``init_Fk_DH_angle`` and the
    GUI) also run on the GPU kernel with N=1 and return float32 numpy (32,3) like the reference.
  * There is no CPU path: without a CUDA device every call raises.
"""
from __future__ import annotations

import numpy as np
import torch

from . import tables
from .functional import fk_world16, fk_world16_wide, wide_rows_ok
from .functional import scatter_16_to_32 as _scatter32
from .tables import used_16key_15bone_len_table  # noqa: F401  (re-exported like the reference module)

# angle ranges of the non-GAN sampler (forward_kinematics_DH_model.py:935-976): joint1..joint34, degrees;
# row 24 ('joint24': {}) is skipped by the sampler, kept here as (0, 0)
SAMPLER_ANGLE_RANGE = (
    (-90, 45), (-90, 45), (-45, 120), (-135, 0), (0, 0), (-45, 90), (-45, 90), (-45, 120), (-135, 0), (0, 0),
    (-25, 25), (-10, 90), (-20, 20), (-20, 20), (-10, 45), (-25, 25), (-20, 20), (0, 0), (-20, 20), (-90, 90),
    (-20, 90), (-45, 45), (0, 0), (0, 0), (-135, 45), (-135, 45), (-45, 180), (0, 135), (0, 0), (-45, 135),
    (-45, 135), (-45, 180), (0, 135), (0, 0))
SAMPLER_GLOBAL_ROT_RANGE = ((-20, 20), (-20, 20), (-180, 180))

H36M_POINTS_LEFT = [6, 7, 8, 17, 18, 19]
H36M_POINTS_RIGHT = [1, 2, 3, 25, 26, 27]

_BONE_KWARGS = tables.BONE_NAMES
_IDX16 = None


def _index16(device):
    global _IDX16
    if _IDX16 is None or _IDX16.device != device:
        _IDX16 = torch.as_tensor(tables.H36M_32_To_16_Table, dtype=torch.long, device=device)
    return _IDX16


def _col(v, like):
    """[N] float32 tensor on `like`'s device from a tensor, a Python number or a numpy value."""
    if torch.is_tensor(v):
        return v.to(device=like.device, dtype=torch.float32).reshape(-1).expand(like.shape[0]) if v.numel() == 1 \
            else v.to(device=like.device, dtype=torch.float32).reshape(-1)
    return torch.full((like.shape[0],), float(v), dtype=torch.float32, device=like.device)


def dh_matrix(alpha, a, d, theta, args=None):
    """One modified-DH 4x4 transform per row, degrees in (forward_kinematics_DH_model.py:53-116).
    Scalar inputs: float64 numpy (4,4), the reference's numpy branch (:54-78; the GUI uses it).  Tensor `theta` [N]:
    float32 [N,4,4] on theta's device, differentiable -- the same 16 entries the reference writes one column at a time
    into a pre-sized buffer (:99-114), here N comes from theta, not from args.batch_size.  The fused kernels never
    build these matrices (every alpha is 0 or +-90 deg, so a joint step is a signed axis permutation and one planar
    rotation); this function exists for callers that want the matrices themselves."""
    if torch.is_tensor(theta):
        th = theta.reshape(-1).to(torch.float32)
        al = _col(alpha, th) / 180 * np.pi
        th = th / 180 * torch.tensor(np.pi, dtype=torch.float32, device=th.device)     # as the reference rounds it (:91)
        a_, d_ = _col(a, th), _col(d, th)
        ct, st, ca, sa = torch.cos(th), torch.sin(th), torch.cos(al), torch.sin(al)
        z, o = torch.zeros_like(th), torch.ones_like(th)
        rows = [ct, -st, z, a_, st * ca, ct * ca, -sa, -sa * d_, st * sa, ct * sa, ca, ca * d_, z, z, z, o]
        return torch.stack(rows, dim=1).view(-1, 4, 4)
    al, th = alpha / 180 * np.pi, theta / 180 * np.pi
    ca, sa, ct, st = np.cos(al), np.sin(al), np.cos(th), np.sin(th)
    return np.array([[ct, -st, 0.0, a], [st * ca, ct * ca, -sa, -sa * d], [st * sa, ct * sa, ca, ca * d],
                     [0.0, 0.0, 0.0, 1.0]])


def rotationMatrix(angle_x, angle_y, angle_z, args=None):
    """R = Rx(angle_x) Ry(angle_y) Rz(angle_z), degrees in (forward_kinematics_DH_model.py:118-191).  Scalars: numpy
    (3,3); tensors [N]: float32 [N,3,3] on their device, differentiable (the closed form of the reference's
    R1.bmm(R2).bmm(R3), the one the kernels' `global_rotation` evaluates)."""
    if torch.is_tensor(angle_x):
        ax = angle_x.reshape(-1).to(torch.float32) / 180 * np.pi
        ay, az = _col(angle_y, ax) / 180 * np.pi, _col(angle_z, ax) / 180 * np.pi
        sx, cx, sy, cy, sz, cz = torch.sin(ax), torch.cos(ax), torch.sin(ay), torch.cos(ay), torch.sin(az), torch.cos(az)
        rows = [cy * cz, -cy * sz, sy,
                sx * sy * cz + cx * sz, -sx * sy * sz + cx * cz, -sx * cy,
                -cx * sy * cz + sx * sz, cx * sy * sz + sx * cz, cx * cy]
        return torch.stack(rows, dim=1).view(-1, 3, 3)
    ax, ay, az = (v / 180 * np.pi for v in (angle_x, angle_y, angle_z))
    r1 = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    r2 = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    r3 = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    return r1.dot(r2).dot(r3)


def _shared_angle_view(rleg, lleg, body, rhand, lhand):
    """The generator passes the five angle kwargs as column slices of ONE [N,37] tensor (Fk_generator.py:179-184).
    When that is the case return the [N,33] view of their common base in kernel order (right leg, left leg, body, right
    hand, left hand) -- no torch.cat, no split in the backward, the kernel reads the base tensor with its row stride.
    Returns None when the five tensors are anything else."""
    parts = (rleg, lleg, body, rhand, lhand)
    base = rleg._base
    if base is None or base.dim() != 2 or base.stride(1) != 1 or any(t._base is not base for t in parts):
        return None
    n, stride = base.shape[0], base.stride(0)
    col0 = rleg.storage_offset() - base.storage_offset()
    if col0 < 0 or col0 + 33 > base.shape[1]:
        return None
    for t, off, width in zip(parts, (0, 5, 10, 23, 28), (5, 5, 13, 5, 5)):
        if (t.dim() != 2 or t.shape != (n, width) or t.stride(1) != 1 or (n > 1 and t.stride(0) != stride)
                or t.storage_offset() - base.storage_offset() != col0 + off):
            return None
    return base[:, col0:col0 + 33]


def _wide_layout(ang_view, grot):
    """(base, goff) when the angle view starts at column 0 of a base tensor that also holds the global rotation (the
    generator's [N,37] slot tensor, Fk_generator.py:136-184) and that tensor can go to the kernels as it is; else None."""
    base = ang_view._base
    if base is None or grot._base is not base or ang_view.storage_offset() != base.storage_offset():
        return None
    goff = grot.storage_offset() - base.storage_offset()
    if grot.dim() != 2 or grot.shape != (base.shape[0], 3) or grot.stride(1) != 1 or \
            (base.shape[0] > 1 and grot.stride(0) != base.stride(0)):
        return None
    return (base, goff) if wide_rows_ok(base, goff) else None


class LazyWorld32:
    """The [N,32,3] tensor `change_3d_joint_angle` returns, not yet materialised ("lazy 32-slot scatter only
    when a caller really indexes it").

    Every tensor-branch caller of the reference immediately gathers the 16 used joints back out,
    ``x[:, H36M_32_To_16_Table]`` (Fk_generator.py:259, :453) -- with the 32-slot layout materialised that is the FK
    kernel, a scatter kernel, a torch index gather and, in the backward, an index_put into zeros plus the scatter's
    transpose: 5 launches and ~2.4 kB per pose where the FK kernel alone moves 408 B.  This object answers exactly that
    index with the kernel's own [N,16,3] output (autograd history intact) and turns into the real [N,32,3] tensor --
    one `dhfk_scatter32_forward` launch -- on any other use: any other index, any attribute or method of
    torch.Tensor, any torch.* function it is passed to (`__torch_function__`).  It is not a torch.Tensor instance;
    `.tensor()` returns one."""

    __slots__ = ("_w16", "_root", "_t32")

    def __init__(self, world16, root):
        self._w16, self._root, self._t32 = world16, root, None

    # ---- the cheap answers ------------------------------------------------------------------------------
    @property
    def shape(self):
        return torch.Size((self._w16.shape[0], 32, 3))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 3

    @property
    def device(self):
        return self._w16.device

    @property
    def dtype(self):
        return self._w16.dtype

    @property
    def is_cuda(self):
        return self._w16.is_cuda

    @property
    def requires_grad(self):
        return self._w16.requires_grad or self._root.requires_grad

    def __len__(self):
        return self._w16.shape[0]

    @staticmethod
    def _is_joint_table(idx):
        if isinstance(idx, torch.Tensor):
            return idx.dim() == 1 and idx.numel() == 16 and not idx.is_cuda and idx.tolist() == tables.H36M_32_To_16_Table
        if isinstance(idx, np.ndarray):
            return idx.shape == (16,) and idx.tolist() == tables.H36M_32_To_16_Table
        return isinstance(idx, (list, tuple)) and list(idx) == tables.H36M_32_To_16_Table

    def __getitem__(self, index):
        if (isinstance(index, tuple) and len(index) == 2 and isinstance(index[0], slice) and index[0] == slice(None)
                and self._is_joint_table(index[1])):
            return self._w16                       # x[:, H36M_32_To_16_Table]: the kernel's own output, no scatter
        return self.tensor()[index]

    # ---- everything else: the real tensor ----------------------------------------------------------------
    def tensor(self):
        if self._t32 is None:
            self._t32 = _scatter32(self._w16, self._root.reshape(self._w16.shape[0], 3))
        return self._t32

    def __getattr__(self, name):                   # only reached for names not defined above
        return getattr(self.tensor(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        unwrap = lambda a: a.tensor() if isinstance(a, LazyWorld32) else a
        args = tuple(unwrap(a) if not isinstance(a, (list, tuple)) else type(a)(unwrap(b) for b in a) for a in args)
        kwargs = {k: unwrap(v) for k, v in (kwargs or {}).items()}
        return func(*args, **kwargs)

    def __array__(self, dtype=None):
        a = self.tensor().detach().cpu().numpy()
        return a if dtype is None else a.astype(dtype)

    def __repr__(self):
        return "LazyWorld32(n=%d, materialised=%s)" % (self._w16.shape[0], self._t32 is not None)


def _binop(name):
    def op(self, other):
        return getattr(self.tensor(), name)(other.tensor() if isinstance(other, LazyWorld32) else other)
    op.__name__ = name
    return op


for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__", "__neg__",
           "__matmul__", "__pow__", "__eq__", "__ne__", "__lt__", "__le__", "__gt__", "__ge__"):
    if _n == "__neg__":
        LazyWorld32.__neg__ = lambda self: -self.tensor()
    else:
        setattr(LazyWorld32, _n, _binop(_n))
LazyWorld32.__hash__ = object.__hash__


def scatter_16_to_32(world16: torch.Tensor, root: torch.Tensor) -> torch.Tensor:
    """[N,16,3] -> the reference's [N,32,3] layout (:745-820): gathered slots hold the joints, slot 14
    duplicates the head joint (slot 15), every other slot equals root (0 + root).  One launch each way
    (dhfk_scatter32_*) instead of expand + clone + index_copy + slice write and their autograd mirror."""
    return _scatter32(world16, root.reshape(world16.shape[0], 3))


class Forward_Kinematics_DH_Model:
    def __init__(self, args, train_subjects, dataset):
        self.args = args
        self.train_subjects = train_subjects
        self.GAN_BATCH_SIZE = getattr(args, "batch_size", None)
        self.dataset = dataset
        self.random = np.random.RandomState(getattr(args, "random_seed", 0))  # Fk_generator.py:201,383 use it

        self.generator_3d_pos_angle = []
        self.generator_global_rot_3d_pos_angle = []
        self.generator_bone_len = []
        self.generator_root = []
        self.show_3d_pos_num = 0
        self.record_bone_len = []
        self.camera_parameters = {}
        self.root_3d_pos = np.array([0, 0, 0])
        self.dataSet_world_3d_pos = {}
        self.dataSet_2d_pos = {}
        self.choice_subject = []
        self.choice_action = []
        self.choice_cam = []

        self.right_leg_joint_num = 5
        self.left_leg_joint_num = 5
        self.body_joint_num = 13
        self.right_hand_joint_num = 5 + self.body_joint_num - 4
        self.left_hand_joint_num = 5 + self.body_joint_num - 4
        # the constant DH tables as plain lists, for callers that read them (GUI); degrees / metres
        a = tables.ALPHA_DEG.tolist()
        t0 = tables.THETA0_DEG.tolist()
        self.right_leg_joints_alpha, self.right_leg_joints_theta = a[0:5], t0[0:5]
        self.left_leg_joints_alpha, self.left_leg_joints_theta = a[5:10], t0[5:10]
        self.body_joints_alpha, self.body_joints_theta = a[10:23], t0[10:23]
        self.right_hand_joints_alpha, self.right_hand_joints_theta = a[23:28], t0[23:28]
        self.left_hand_joints_alpha, self.left_hand_joints_theta = a[28:33], t0[28:33]

        self.real_used_num = 1
        if getattr(args, "single_or_multi_train_mode", "single") == "multi":
            frames = 1
            for f in [int(x) for x in args.architecture.split(",")]:
                frames *= f  # video_receptive_field = product (video_mode_operate.py:411-415)
            self.real_used_num = frames

    # ------------------------------------------------------------------------------------------
    def change_3d_joint_angle(self, left_leg_joints_angle, right_leg_joints_angle, body_joints_angle,
                              left_hand_joints_angle, right_hand_joints_angle, generator_global_rot_3d_pos_angle,
                              left_small_leg_len, right_small_leg_len, left_big_leg_len, right_big_leg_len,
                              left_hip_len, right_hip_len, waist_len, thorax_len, left_shoulder_len,
                              right_shoulder_len, left_big_arm_len, right_big_arm_len, left_small_arm_len,
                              right_small_arm_len, neck_len, root_3d_pos):
        lens = (left_small_leg_len, right_small_leg_len, left_big_leg_len, right_big_leg_len, left_hip_len,
                right_hip_len, waist_len, thorax_len, left_shoulder_len, right_shoulder_len, left_big_arm_len,
                right_big_arm_len, left_small_arm_len, right_small_arm_len, neck_len)
        if not torch.is_tensor(left_leg_joints_angle):
            return self._single_pose_numpy(left_leg_joints_angle, right_leg_joints_angle, body_joints_angle,
                                           left_hand_joints_angle, right_hand_joints_angle,
                                           generator_global_rot_3d_pos_angle, lens, root_3d_pos)
        if not torch.cuda.is_available():
            raise RuntimeError("dhfk needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = right_leg_joints_angle.device
        if dev.type != "cuda":
            dev = torch.device("cuda", torch.cuda.current_device())
        to = lambda t: t if t.device == dev else t.to(dev)
        # zero-copy view of the generator's [N,37] tensor: the kernel reads the 148-byte rows through its cp.async
        # gather path (0.100 ms per 1 M poses; torch.cat + the packed path: 0.04 + 0.066 ms and one more launch)
        ang = _shared_angle_view(right_leg_joints_angle, left_leg_joints_angle, body_joints_angle,
                                 right_hand_joints_angle, left_hand_joints_angle)
        if ang is None or ang.device != dev:
            ang = torch.cat([to(right_leg_joints_angle), to(left_leg_joints_angle), to(body_joints_angle),
                             to(right_hand_joints_angle), to(left_hand_joints_angle)], dim=1)
        n = ang.shape[0]
        grot = to(generator_global_rot_3d_pos_angle)
        if all(type(v) is torch.Tensor and v.dim() == 1 and v.device == dev for v in lens):
            bone = torch.stack(lens, dim=1)           # the generator's 15 length products: one launch, no per-column glue
        else:
            cols = []
            for v in lens:
                if not torch.is_tensor(v):
                    v = torch.full((n,), float(v), dtype=torch.float32, device=dev)
                cols.append(to(v).reshape(-1))
            bone = torch.stack(cols, dim=1)
        root = to(root_3d_pos).reshape(-1, 3)
        # forward_kinematics_DH_model.py:564 keeps this attribute; a detached view, so that the model does not keep the
        # previous iteration's autograd graph (and the leaves' AccumulateGrad nodes, bound to the stream of their first
        # use) alive -- that is what breaks a later CUDA-graph capture of the caller's step
        self.global_rot_angle = (generator_global_rot_3d_pos_angle.detach()
                                 if torch.is_tensor(generator_global_rot_3d_pos_angle) else generator_global_rot_3d_pos_angle)
        wide = _wide_layout(ang, grot) if ang._base is not None else None
        if wide is not None:      # the generator's own [N,37] tensor goes to the kernels as one slab per tile
            world16 = fk_world16_wide(wide[0], wide[1], bone, root)
        else:
            world16 = fk_world16(ang, grot, bone, root)
        # the reference keeps its result as an attribute (forward_kinematics_DH_model.py:745-820); here the attribute is a
        # detached twin of what is returned, for the same reason as global_rot_angle above
        self.single_generator_3d_world_32keyPoint = LazyWorld32(world16.detach(), root.detach())
        return LazyWorld32(world16, root)

    def _single_pose_numpy(self, ll, rl, body, lh, rh, grot, lens, root):
        if not torch.cuda.is_available():
            raise RuntimeError("dhfk needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
        ang = np.concatenate([np.asarray(rl, np.float64), np.asarray(ll, np.float64), np.asarray(body, np.float64),
                              np.asarray(rh, np.float64), np.asarray(lh, np.float64)]).astype(np.float32)
        f = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32).reshape(1, -1), device=dev)
        with torch.no_grad():
            world16 = fk_world16(f(ang), f(grot), f(lens), f(root))
            out = scatter_16_to_32(world16, f(root))
        self.single_generator_3d_world_32keyPoint = out[0].cpu().numpy().astype(np.float32)
        return self.single_generator_3d_world_32keyPoint

    def init_Fk_DH_angle(self):
        """T-pose known answer of the reference (:824-858)."""
        return self.change_3d_joint_angle(
            left_leg_joints_angle=[0] * 5, right_leg_joints_angle=[0] * 5, body_joints_angle=[0] * 13,
            left_hand_joints_angle=[0] * 5, right_hand_joints_angle=[0] * 5,
            generator_global_rot_3d_pos_angle=(0.0, 0.0, 0.0),
            left_small_leg_len=0.5, right_small_leg_len=0.5, left_big_leg_len=0.6, right_big_leg_len=0.6,
            left_hip_len=0.25, right_hip_len=0.25, waist_len=0.25, thorax_len=0.2, left_shoulder_len=0.4,
            right_shoulder_len=0.4, left_big_arm_len=0.4, right_big_arm_len=0.4, left_small_arm_len=0.35,
            right_small_arm_len=0.35, neck_len=0.15, root_3d_pos=(0.0, 0.0, 0.0))

    # ---- host-side bookkeeping kept for interface compatibility (:861-929) ---------------------
    def random_state(self):
        return self.random

    def set_random_state(self, random):
        self.random = random

    def get_dataSet_3d_and_2d_pose(self, dataset, pix_2d):
        world_3d = {}
        for subject in dataset.subjects():
            world_3d[subject] = {}
            for action in dataset[subject].keys():
                anim = dataset[subject][action]
                world_3d[subject][action] = {cam_idx: anim["positions"]
                                             for cam_idx, _ in enumerate(pix_2d[subject][action])}
        self.dataSet_world_3d_pos = world_3d
        self.dataSet_2d_pos = pix_2d

    def my_random_get_sigle_frame_data(self):
        # RNG draw order must match the reference (:883-898) so seeded runs pick the same frames
        subject = self.train_subjects[self.random.randint(0, len(self.train_subjects))]
        actions = list(self.dataSet_world_3d_pos[subject].keys())
        action = actions[self.random.randint(0, len(actions))]
        cams = list(self.dataSet_world_3d_pos[subject][action].keys())
        cam = cams[self.random.randint(0, len(cams))]
        frame = self.random.randint(0, self.dataSet_world_3d_pos[subject][action][cam].shape[0])
        return subject, action, cam, frame

    def get_bone_len_from_dataSet(self):
        subject, action, cam, frame = self.my_random_get_sigle_frame_data()
        pose = np.array(self.dataSet_world_3d_pos[subject][action][cam][frame], copy=True)
        self.record_bone_len = [float(np.linalg.norm(pose[i] - pose[j])) for (i, j) in used_16key_15bone_len_table]

    def get_root_3d_pos_from_dataSet(self):
        subject, action, cam, frame = self.my_random_get_sigle_frame_data()
        pose = np.array(self.dataSet_world_3d_pos[subject][action][cam][frame], copy=True)
        self.root_3d_pos = pose[0].copy()

    # ---- non-GAN augmentation sampler (--data_enhancement_method normal), :931-1152 ---------------------
    def sample_normal_mode(self):
        """Host-side draws of handler_but_generater in the reference's exact RNG order (so a seeded run
        produces the same poses): returns (angles [W,33], global_rot [W,3], bone_len [W,15] incl. scaler,
        root [W,3]) as float64 numpy, plus the raw lists the reference returns."""
        args = self.args
        W = int(args.generator_whole_number)
        n_tab = len(SAMPLER_ANGLE_RANGE)     # 34
        self.generator_3d_pos_angle, self.generator_global_rot_3d_pos_angle = [], []
        self.generator_bone_len, self.generator_root = [], []
        for frame in range(W):
            if args.generator_choose_BoneLen:
                self.get_bone_len_from_dataSet()
            self.generator_bone_len.append(self.record_bone_len)
            if args.generator_choose_root_pos:
                self.get_root_3d_pos_from_dataSet()
            self.generator_root.append(self.root_3d_pos)
            n_change = self.random.randint(0, n_tab)
            change = self.random.choice(np.arange(n_tab), size=n_change, replace=False)
            row = []
            for j in range(n_tab):
                if j + 1 == 24:
                    continue
                if (j in change) and frame > 0:
                    lo, hi = SAMPLER_ANGLE_RANGE[j]
                    row.append(min(max(self.random.normal((lo + hi) / 2, 60), lo), hi))
                else:
                    row.append(0)
            g = []
            for lo, hi in SAMPLER_GLOBAL_ROT_RANGE:
                if frame > 0 and args.generator_global_rot:
                    g.append(min(max(self.random.normal((lo + hi) / 2, 60), lo), hi))
                else:
                    g.append(0)
            self.generator_global_rot_3d_pos_angle.append(g)
            self.generator_3d_pos_angle.append(row)
        self.generator_3d_pos_angle = np.array(self.generator_3d_pos_angle).reshape(-1, n_tab - 1)
        bone = np.zeros((W, 15))
        grp = tables.BONE_SCALER_GROUP
        for frame in range(W):
            scaler = np.zeros(8)
            if args.bone_len_scaler == "different":
                scaler = np.array(self.random.randint(-200, 200, size=(8))).reshape(8) / 1000.0
            elif args.bone_len_scaler == "same":
                scaler = np.array(self.random.randint(-200, 200, size=(1))).reshape(1).repeat(8) / 1000.0
            elif args.bone_len_scaler == "":
                pass
            else:
                raise ValueError("args.bone_len_scaler")
            base = np.asarray(self.generator_bone_len[frame], dtype=np.float64)
            bone[frame] = base * np.where(grp >= 0, 1 + scaler[np.maximum(grp, 0)], 1.0)
        root = np.asarray(self.generator_root, dtype=np.float64).reshape(W, 3)
        glob = np.asarray(self.generator_global_rot_3d_pos_angle, dtype=np.float64).reshape(W, 3)
        return self.generator_3d_pos_angle.astype(np.float64), glob, bone, root

    def handler_but_generater(self):
        """Same return tuple as the reference; the W forward-kinematics evaluations the reference does one
        pose at a time in numpy (:1071-1144) are one fused kernel launch here."""
        ang, glob, bone, root = self.sample_normal_mode()
        if not torch.cuda.is_available():
            raise RuntimeError("dhfk needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
        f = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=dev)
        with torch.no_grad():
            r = f(root)
            pos32 = scatter_16_to_32(fk_world16(f(ang), f(glob), f(bone), r), r)
        self.record_bone_len = self.generator_bone_len[-1] if self.generator_bone_len else self.record_bone_len
        self.root_3d_pos = self.generator_root[-1] if self.generator_root else self.root_3d_pos
        return (pos32.cpu().numpy().astype(np.float32), self.generator_3d_pos_angle,
                self.generator_global_rot_3d_pos_angle, self.generator_bone_len, self.generator_root)
