"""Synthetic inputs for parity tests and benchmarks.  Host-side numpy RNG with explicit
seeds so the CPU oracle, the golden fixtures and the GPU kernels all see identical bits."""
from __future__ import annotations

import numpy as np

from . import tables


def gan_like(n: int, seed: int = 1234, root_mode: str = "volume", angle_mode: str = "gan"):
    """-> dict(ang [n,33] deg, grot [n,3] deg, bone [n,15] m, root [n,3] m), float32 numpy.

    angle_mode 'gan': u ~ U(-1,1) per slot mapped through GAN_angle_range_table rows *as the generator
      applies them* (slot i <-> 'joint{i+1}', Fk_generator.py:143-151, including the one-row shift of
      the arm chains) ; 'stress': U(-180,180) on every slot.
    grot: U(-180,180) (Fk_generator.py:35-39).
    bone: a random row of the S1,5,6,7,8 templates in used_16key_15bone_len_table order times
      (1 + s/1000), s ~ integer U[-200,200) per symmetric group, thorax unscaled (Fk_generator.py:196-230).
    root_mode 'volume': U(-1,1) x U(-1,1) x U(0.8,1.2) (keeps |x/z| < 1 for the H36M cameras);
      'generator': 10*tanh(randn) (Fk_generator.py:122; the projection clamp becomes active).
    """
    rng = np.random.RandomState(seed)
    u = rng.uniform(-1.0, 1.0, size=(n, 33)).astype(np.float32)
    if angle_mode == "gan":
        lo = tables.GAN_ANGLE_RANGE[:33, 0]
        hi = tables.GAN_ANGLE_RANGE[:33, 1]
        ang = u * ((hi - lo) / 2) + (hi + lo) / 2
        ang[:, [4, 9, 22, 23, 28]] = 0.0  # slots the generator pins to zero (Fk_generator.py:136)
    elif angle_mode == "stress":
        ang = u * 180.0
    else:
        raise ValueError(angle_mode)
    grot = rng.uniform(-180.0, 180.0, size=(n, 3)).astype(np.float32)
    rows = rng.randint(0, tables.BONE_TEMPLATES.shape[0], size=n)
    scal = rng.randint(-200, 200, size=(n, 8)).astype(np.float32) / 1000.0
    grp = tables.BONE_SCALER_GROUP
    factor = np.where(grp[None, :] >= 0, 1.0 + scal[:, np.maximum(grp, 0)], 1.0).astype(np.float32)
    bone = (tables.BONE_TEMPLATES[rows] * factor).astype(np.float32)
    if root_mode == "volume":
        root = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(0.8, 1.2, n)], 1)
    elif root_mode == "generator":
        root = 10.0 * np.tanh(rng.randn(n, 3))
    else:
        raise ValueError(root_mode)
    return dict(ang=np.ascontiguousarray(ang, np.float32), grot=grot, bone=bone,
                root=np.ascontiguousarray(root, np.float32))


def upstream_grads(n: int, seed: int = 4321):
    rng = np.random.RandomState(seed)
    return dict(g_world=rng.randn(n, 16, 3).astype(np.float32), g_cam=rng.randn(n, 16, 3).astype(np.float32),
                g_uv=rng.randn(n, 16, 2).astype(np.float32))


def gan_like_torch(n: int, device, seed: int = 1234, root_mode: str = "volume", angle_mode: str = "gan"):
    """Same distributions drawn directly on `device` with a torch generator (for 1M+ pose benchmarks
    where a host draw + H2D would dominate set-up time).  Not bit-identical to gan_like()."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    u = torch.rand((n, 33), generator=g, device=device) * 2 - 1
    if angle_mode == "gan":
        lo = torch.as_tensor(tables.GAN_ANGLE_RANGE[:33, 0], device=device)
        hi = torch.as_tensor(tables.GAN_ANGLE_RANGE[:33, 1], device=device)
        ang = u * ((hi - lo) / 2) + (hi + lo) / 2
        ang[:, [4, 9, 22, 23, 28]] = 0.0
    else:
        ang = u * 180.0
    grot = torch.rand((n, 3), generator=g, device=device) * 360 - 180
    tmpl = torch.as_tensor(tables.BONE_TEMPLATES, device=device)
    rows = torch.randint(0, tmpl.shape[0], (n,), generator=g, device=device)
    scal = torch.randint(-200, 200, (n, 8), generator=g, device=device).float() / 1000.0
    grp = torch.as_tensor(tables.BONE_SCALER_GROUP, device=device)
    factor = torch.where(grp[None, :] >= 0, 1.0 + scal[:, grp.clamp(min=0)], torch.ones((), device=device))
    bone = tmpl[rows] * factor
    if root_mode == "volume":
        r = torch.rand((n, 3), generator=g, device=device)
        root = torch.stack([r[:, 0] * 2 - 1, r[:, 1] * 2 - 1, r[:, 2] * 0.4 + 0.8], 1)
    else:
        root = 10.0 * torch.tanh(torch.randn((n, 3), generator=g, device=device))
    return dict(ang=ang.contiguous(), grot=grot.contiguous(), bone=bone.contiguous(), root=root.contiguous())
