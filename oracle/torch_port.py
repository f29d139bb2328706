"""TEST / BASELINE INFRASTRUCTURE ONLY -- torch-CPU restatement of the reference's own
FK + camera + projection path, issuing the SAME torch op sequence the reference issues
(zeros + strided slice assignment per DH matrix entry, clone+bmm chain products, per-chain
global-rotation bmm, 51 scalar-column scatters, qrot by two torch.cross, clamp/sum/cat
projection) so that (a) its results are bit-identical to the reference on the same host and
(b) its run time is the reference's CPU run time.  `/root/reference` cannot travel to the
GPU box; this file can.  It is what ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs time (kind = "port").

Pinned by tests/test_oracle_golden.py: outputs and autograd gradients equal the golden
vectors produced by the unmodified reference (oracle/make_golden.py) -- bit-exact on the
host that generated them, 1e-6 elsewhere.

Never imported by the product package.  Citations are relative to
/root/reference/DH-AUG_master.
"""
from __future__ import annotations

import numpy as np
import torch

# DH constant tables, models_Fk_GAN/forward_kinematics_DH_model.py:234-261.
# chain order used here: rleg, lleg, body, rhand, lhand
_ALPHA = (
    (0.0, -90.0, -90.0, 0.0, 0.0),
    (0.0, 90.0, 90.0, 0.0, 0.0),
    (0.0,) + (-90.0,) * 11 + (90.0,),
    (-90.0, -90.0, -90.0, 0.0, 0.0),
    (-90.0, 90.0, 90.0, 0.0, 0.0),
)
_A0 = (
    (0.25, 0.0, 0.0, 0.6, 0.5),
    (-0.25, 0.0, 0.0, 0.6, 0.5),
    (0.0,) * 12 + (0.15,),
    (-0.3, 0.0, 0.0, 0.4, 0.35),
    (0.3, 0.0, 0.0, 0.4, 0.35),
)
_D0 = (
    (0.0,) * 5,
    (0.0,) * 5,
    (0.0, 0.0, 0.0, 0.25, 0.0, 0.0, 0.2) + (0.0,) * 6,
    (0.0,) * 5,
    (0.0,) * 5,
)
_THETA0 = (
    (0.0, -90.0, 180.0, 0.0, 0.0),
    (180.0, -90.0, 0.0, 0.0, 0.0),
    (90.0,) + (-90.0,) * 10 + (0.0, 0.0),
    (-180.0, -90.0, 180.0, 0.0, 0.0),
    (0.0, -90.0, 0.0, 0.0, 0.0),
)
_ANG_SLICE = ((0, 5), (5, 10), (10, 23), (23, 28), (28, 33))  # Fk_generator.py:179-184
# (chain, index, 'a'|'d', bone index, sign): forward_kinematics_DH_model.py:571-589
_LEN_WRITES = (
    (1, 0, "a", 4, -1), (1, 3, "a", 2, 1), (1, 4, "a", 0, 1),
    (0, 0, "a", 5, 1), (0, 3, "a", 3, 1), (0, 4, "a", 1, 1),
    (2, 12, "a", 14, 1), (2, 3, "d", 6, 1), (2, 6, "d", 7, 1),
    (4, 0, "a", 8, 1), (4, 3, "a", 10, 1), (4, 4, "a", 12, 1),
    (3, 0, "a", 9, -1), (3, 3, "a", 11, 1), (3, 4, "a", 13, 1),
)
# 32-slot scatter (slot, product chain, index): forward_kinematics_DH_model.py:751-817
_SCATTER = (
    (0, 2, 0), (1, 0, 0), (2, 0, 3), (3, 0, 4), (6, 1, 0), (7, 1, 3), (8, 1, 4),
    (12, 2, 3), (13, 2, 6), (14, 2, 12), (15, 2, 12), (17, 4, 9), (18, 4, 12), (19, 4, 13),
    (25, 3, 9), (26, 3, 12), (27, 3, 13),
)
H36M_32_TO_16 = [0, 1, 2, 3, 6, 7, 8, 12, 13, 15, 17, 18, 19, 25, 26, 27]  # common/h36m_dataset.py:37-38


class RefFKPort:
    """Stateful like the reference object (tables replicated to [N,n]; lengths written in place)."""

    def __init__(self, n: int, device=None):
        self.n = n
        # device=None / cpu: the reference's CPU branch (what the parity pin and the CPU baseline use).  A CUDA device
        # runs the same op sequence in torch eager on the GPU with every tensor created in place -- a generous
        # stand-in for the reference's own CUDA branch, which builds each matrix on the host and copies it over
        # (forward_kinematics_DH_model.py:93-97,154-156); used only as context by bench.py.
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        rep = lambda rows: torch.tensor(np.tile(np.asarray(rows, dtype=np.float64), (n, 1)), dtype=torch.float32).to(self.device)
        self.alpha = [rep(r) for r in _ALPHA]
        self.a = [rep(r) for r in _A0]
        self.d = [rep(r) for r in _D0]
        self.theta0 = [rep(r) for r in _THETA0]

    # forward_kinematics_DH_model.py:80-116 (torch branch of dh_matrix)
    def _dh(self, alpha, a, d, theta):
        alpha = alpha / 180 * np.pi
        theta = theta / 180 * torch.tensor(np.pi, dtype=torch.float32, device=self.device)
        m = torch.zeros((self.n, 4, 4), dtype=torch.float32, device=self.device)
        m[:, 0, 0] = torch.cos(theta)
        m[:, 0, 1] = -torch.sin(theta)
        m[:, 0, 2] = 0
        m[:, 0, 3] = a
        m[:, 1, 0] = torch.sin(theta) * torch.cos(alpha)
        m[:, 1, 1] = torch.cos(theta) * torch.cos(alpha)
        m[:, 1, 2] = -torch.sin(alpha)
        m[:, 1, 3] = -torch.sin(alpha) * d
        m[:, 2, 0] = torch.sin(theta) * torch.sin(alpha)
        m[:, 2, 1] = torch.cos(theta) * torch.sin(alpha)
        m[:, 2, 2] = torch.cos(alpha)
        m[:, 2, 3] = torch.cos(alpha) * d
        m[:, 3, 0] = 0
        m[:, 3, 1] = 0
        m[:, 3, 2] = 0
        m[:, 3, 3] = 1
        return m

    # forward_kinematics_DH_model.py:141-191
    def _global_rotation(self, ax, ay, az):
        ax = ax / 180 * np.pi
        ay = ay / 180 * np.pi
        az = az / 180 * np.pi
        n = self.n
        r1 = torch.zeros((n, 3, 3), dtype=torch.float32, device=self.device)
        r2 = torch.zeros((n, 3, 3), dtype=torch.float32, device=self.device)
        r3 = torch.zeros((n, 3, 3), dtype=torch.float32, device=self.device)
        r1[:, 0:] = torch.tensor([1, 0, 0], dtype=torch.float32, device=self.device)
        r1[:, 1, 0] = 0
        r1[:, 1, 1] = torch.cos(ax)
        r1[:, 1, 2] = -torch.sin(ax)
        r1[:, 2, 0] = 0
        r1[:, 2, 1] = torch.sin(ax)
        r1[:, 2, 2] = torch.cos(ax)
        r2[:, 0, 0] = torch.cos(ay)
        r2[:, 0, 1] = 0
        r2[:, 0, 2] = torch.sin(ay)
        r2[:, 1:] = torch.tensor([0, 1, 0], dtype=torch.float32, device=self.device)
        r2[:, 2, 0] = -torch.sin(ay)
        r2[:, 2, 1] = 0
        r2[:, 2, 2] = torch.cos(ay)
        r3[:, 0, 0] = torch.cos(az)
        r3[:, 0, 1] = -torch.sin(az)
        r3[:, 0, 2] = 0
        r3[:, 1, 0] = torch.sin(az)
        r3[:, 1, 1] = torch.cos(az)
        r3[:, 1, 2] = 0
        r3[:, 2:] = torch.tensor([0, 0, 1], dtype=torch.float32, device=self.device)
        return r1.bmm(r2).bmm(r3)

    # forward_kinematics_DH_model.py:562-822
    def fk32(self, angles33, grot3, bone15, root3):
        n = self.n
        rg = self._global_rotation(grot3[:, 0], grot3[:, 1], grot3[:, 2])
        for c, i, kind, b, sign in _LEN_WRITES:
            tgt = self.a if kind == "a" else self.d
            tgt[c][:, i] = bone15[:, b] if sign > 0 else -bone15[:, b]
        # order of construction in the reference: lleg, rleg, body, rhand, lhand
        hm = [None] * 5
        for c in (1, 0, 2):
            lo, hi = _ANG_SLICE[c]
            ang = angles33[:, lo:hi]
            h = torch.zeros((n, hi - lo, 4, 4), dtype=torch.float32, device=self.device)
            for i in range(hi - lo):
                h[:, i] = self._dh(self.alpha[c][:, i], self.a[c][:, i], self.d[c][:, i],
                                   self.theta0[c][:, i] + ang[:, i])
            hm[c] = h
        for c in (3, 4):
            lo, hi = _ANG_SLICE[c]
            ang = angles33[:, lo:hi]
            h = torch.zeros((n, 9 + 5, 4, 4), dtype=torch.float32, device=self.device)
            h[:, 0:9] = torch.clone(hm[2][:, 0:9])
            for i in range(5):
                h[:, i + 9] = self._dh(self.alpha[c][:, i], self.a[c][:, i], self.d[c][:, i],
                                       self.theta0[c][:, i] + ang[:, i])
            hm[c] = h
        for c in (1, 0, 2, 3, 4):
            h = hm[c]
            for i in range(h.shape[1] - 1):
                h[:, i + 1] = torch.bmm(torch.clone(h[:, i]), torch.clone(h[:, i + 1]))
        pos = [None] * 5
        for c in (1, 0, 2, 4, 3):
            h = hm[c]
            x = torch.clone(h[:, :, 0, 3])
            y = torch.clone(h[:, :, 1, 3])
            z = torch.clone(h[:, :, 2, 3])
            p = torch.zeros((n, 3, h.shape[1]), dtype=torch.float32, device=self.device)
            p[:, 0, :] = x[:, :]
            p[:, 1, :] = y[:, :]
            p[:, 2, :] = z[:, :]
            pos[c] = rg.bmm(p)
        out = torch.zeros((n, 32, 3), dtype=torch.float32, device=self.device)
        for slot, c, i in _SCATTER:
            for ax in range(3):
                out[:, slot, ax] = pos[c][:, ax, i]
        return out + root3.view(-1, 1, 3)


# common/quaternion.py:6-35 and common/camera.py:36-38
def world_to_camera(x, q, t):
    qi = torch.cat((q[..., :1], -q[..., 1:]), dim=len(q.shape) - 1)
    qr = qi.repeat(x.shape[:-1] + (1,))
    v = x - t
    qvec = qr[..., 1:]
    uv = torch.cross(qvec, v, dim=len(qr.shape) - 1)
    uuv = torch.cross(qvec, uv, dim=len(qr.shape) - 1)
    return v + 2 * (qr[..., :1] * uv + uuv)


# common/camera.py:62-94
def project_to_2d(x, camera_params):
    assert x.shape[-1] == 3
    assert len(camera_params.shape) == 2
    assert camera_params.shape[-1] == 9 or camera_params.shape[-1] == 16
    assert x.shape[0] == camera_params.shape[0]
    while len(camera_params.shape) < len(x.shape):
        camera_params = camera_params.unsqueeze(1)
    f = camera_params[..., :2]
    c = camera_params[..., 2:4]
    k = camera_params[..., 4:7]
    p = camera_params[..., 7:9]
    xx = torch.clamp(x[..., :2] / x[..., 2:], min=-1, max=1)
    r2 = torch.sum(xx[..., :2] ** 2, dim=len(xx.shape) - 1, keepdim=True)
    radial = 1 + torch.sum(k * torch.cat((r2, r2 ** 2, r2 ** 3), dim=len(r2.shape) - 1),
                           dim=len(r2.shape) - 1, keepdim=True)
    tan = torch.sum(p * xx, dim=len(xx.shape) - 1, keepdim=True)
    xxx = xx * (radial + tan) + p * r2
    return f * xxx + c


def pipeline(angles33, grot3, bone15, root3, cam16):
    """FK -> gather 16 -> world->camera -> project.  cam16 = [q4,t3,f2,c2,k3,p2] (array-like)."""
    n = angles33.shape[0]
    cam16 = torch.as_tensor(np.asarray(cam16, dtype=np.float32)).to(angles33.device)
    w32 = RefFKPort(n, angles33.device).fk32(angles33, grot3, bone15.detach(), root3)
    w16 = w32[:, H36M_32_TO_16]
    q = cam16[0:4].view(1, 4)
    t = cam16[4:7].view(1, 3)
    cp = cam16[7:16].view(1, 9).repeat(n, 1)
    cam = world_to_camera(w16, q, t)
    uv = project_to_2d(cam, cp)
    return w32, w16, cam, uv


def time_pipeline(inputs, cam16, g_world, g_uv, backward=True, repeats=3, warmup=1):
    """Time the port on CPU (all torch intra-op threads).  Returns (best_seconds, n)."""
    import time
    ang, grot, bone, root = (torch.as_tensor(x, dtype=torch.float32).cpu() for x in inputs)
    n = ang.shape[0]
    gw = torch.as_tensor(g_world, dtype=torch.float32).cpu().view(n, 16, 3)
    gu = torch.as_tensor(g_uv, dtype=torch.float32).cpu().view(n, 16, 2)
    best = float("inf")
    for it in range(warmup + repeats):
        a = ang.clone().requires_grad_(backward)
        g = grot.clone().requires_grad_(backward)
        r = root.clone().requires_grad_(backward)
        t0 = time.perf_counter()
        if backward:
            _, w16, _, uv = pipeline(a, g, bone, r, cam16)
            ((w16 * gw).sum() + (uv * gu).sum()).backward()
        else:
            with torch.no_grad():
                pipeline(a, g, bone, r, cam16)
        dt = time.perf_counter() - t0
        if it >= warmup:
            best = min(best, dt)
    return best, n


# ---- critic input transform (context baseline for tools/gan_step_bench.py) -------------------------------------------
# models_Fk_GAN/special_operate.py:513-539 (Ct rows: bone = x[second] - x[first]) and
# models_Fk_GAN/Fk_discriminator.py:36-146 (30 row writes into a [30,N] buffer, then a transpose)
_KCS_BONES = ((5, 6), (2, 3), (4, 5), (1, 2), (0, 4), (0, 1), (0, 7), (7, 8), (8, 10), (8, 13), (10, 11), (13, 14),
              (11, 12), (14, 15), (8, 9))
_KCS_PAIRS = ((0, 2), (1, 3), (2, 4), (3, 5), (4, 5), (4, 6), (5, 6), (6, 7), (7, 14), (7, 8), (7, 9), (8, 10), (9, 11),
              (10, 12), (11, 13))


def special_kcs(pos_16_3d):
    """special_KCS_Input_transform restated with the reference's op sequence: incidence matmul, norms, one row write per
    feature.  Pinned to tests/golden/critic.npz by tests/test_oracle_golden.py."""
    x = pos_16_3d.view(-1, 16, 3)
    ct = torch.zeros(15, 16, device=x.device)
    for b, (i, j) in enumerate(_KCS_BONES):
        ct[b, i] = -1.0
        ct[b, j] = 1.0
    c = ct.transpose(1, 0).repeat([x.size(0), 1, 1]).view(-1, 16, 15)
    bone = torch.matmul(x.permute(0, 2, 1).contiguous(), c).permute(0, 2, 1).transpose(1, 0)   # [15, N, 3]
    length = torch.sqrt(torch.sum(bone ** 2, dim=-1))
    out = torch.zeros((30, x.shape[0]), dtype=torch.float32, device=x.device)
    for q, (i, j) in enumerate(_KCS_PAIRS):
        out[q] = torch.sum(bone[i] * bone[j], dim=-1) / (length[i] * length[j])
    out[15:] = length[0:]
    return out.transpose(0, 1)
