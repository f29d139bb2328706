"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* DH-AUG reference as the parity oracle.

This module is never imported by the product package.  It exists so that
``oracle/make_golden.py`` (run in the build container, where ``/root/reference``
is mounted) can execute the reference's own torch / numpy code and freeze its
outputs as fixtures under ``tests/golden/``.  ``/root/reference`` does not
exist on the GPU box, so nothing that runs there may import this file.

What is wrapped (all paths relative to ``/root/reference/DH-AUG_master``):

* ``models_Fk_GAN/forward_kinematics_DH_model.py:194``  ``Forward_Kinematics_DH_Model``
* ``models_Fk_GAN/forward_kinematics_DH_model.py:354-822`` ``change_3d_joint_angle``
* ``common/camera.py:36-38``  ``GAN_torch_world_to_camera``
* ``common/camera.py:62-94``  ``project_to_2d``
* ``common/h36m_dataset.py:37-38,46-234`` joint table + camera tables

The reference module imports matplotlib (Qt5Agg), pylab, thop, h5py and
tensorboardX at module top; none is used by the FK / projection path, so they
are replaced with dummy modules before the import.  ``pylab`` is special: the
numpy branch of ``dh_matrix`` takes ``cos``/``sin``/``float32`` from
``from pylab import *`` and ``handler_but_generater`` takes ``time`` from it.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

REF_ROOT = os.environ.get("DHFK_REFERENCE_ROOT", "/root/reference/DH-AUG_master")
# the same tree staged as one archive by oracle/stage_ref.py (git-ignored, travels to the GPU box via gpurun)
REF_ZIP = os.environ.get("DHFK_REFERENCE_ZIP") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref",
                                                               "dh_aug_ref.zip")


def reference_available() -> bool:
    """The mounted reference tree (build container)."""
    return os.path.isdir(os.path.join(REF_ROOT, "models_Fk_GAN"))


def staged_reference_available() -> bool:
    return os.path.isfile(REF_ZIP)


def reference_source(prefer_staged: bool = False):
    """sys.path entry the reference is imported from: the mounted tree when present (unless prefer_staged), else the
    staged archive (zipimport), else None."""
    if staged_reference_available() and (prefer_staged or not reference_available()):
        return REF_ZIP
    if reference_available():
        return REF_ROOT
    return None


class _Anything:
    """Attribute sink used for the stubbed plotting modules."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


def _stub_module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    def _attr(attr):
        if attr.startswith("__"):          # inspect.getmodule() walks sys.modules and reads __file__ etc.
            raise AttributeError(attr)
        return _Anything()
    mod.__getattr__ = _attr  # type: ignore[attr-defined]
    sys.modules[name] = mod
    return mod


_IMPORTED = None


def import_reference(force_cpu: bool = True, prefer_staged: bool = False):
    """Return a namespace with the reference's hot-path symbols.  force_cpu=True (default) makes the reference take
    its pure-CPU branch even on a GPU host; force_cpu=False leaves torch.cuda alone (the GPU tests that run the
    reference's own training loops with the drop-in installed)."""
    global _IMPORTED
    if _IMPORTED is not None:
        return _IMPORTED
    src = reference_source(prefer_staged)
    if src is None:
        raise RuntimeError("reference not found: neither %s nor the staged archive %s (python oracle/stage_ref.py)"
                           % (REF_ROOT, REF_ZIP))

    import numpy as np
    import torch

    for name in (
        "matplotlib", "matplotlib.animation", "matplotlib.pyplot", "matplotlib.gridspec",
        "matplotlib.figure", "matplotlib.backends", "matplotlib.backends.backend_qt5agg",
        "matplotlib.patches", "matplotlib.colors", "matplotlib.cm",
        "mpl_toolkits", "mpl_toolkits.mplot3d", "thop", "h5py", "tensorboardX",
    ):
        if name not in sys.modules:
            _stub_module(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None  # type: ignore[attr-defined]
    sys.modules["matplotlib"].__path__ = []  # type: ignore[attr-defined]
    sys.modules["matplotlib.backends"].__path__ = []  # type: ignore[attr-defined]
    sys.modules["mpl_toolkits"].__path__ = []  # type: ignore[attr-defined]
    # `from pylab import *` must provide numpy's namespace (cos, sin, float32, ...) plus `time`.
    import time as _time
    pylab = types.ModuleType("pylab")
    for k in dir(np):
        if not k.startswith("_"):
            setattr(pylab, k, getattr(np, k))
    pylab.time = _time
    pylab.__all__ = [k for k in dir(pylab) if not k.startswith("_")]
    sys.modules["pylab"] = pylab

    # Force the pure-CPU branch of the reference even on a GPU host
    # (the CUDA branch mixes devices inside autograd).
    if force_cpu:
        torch.cuda.is_available = lambda: False  # type: ignore[assignment]

    if src not in sys.path:
        sys.path.insert(0, src)

    from models_Fk_GAN import forward_kinematics_DH_model as fkmod  # noqa: E402
    from common import camera as cammod  # noqa: E402
    from common import h36m_dataset as h36m  # noqa: E402
    from common import quaternion as quatmod  # noqa: E402
    from models_Fk_GAN import Fk_generator as genmod  # noqa: E402

    ns = types.SimpleNamespace(fk=fkmod, camera=cammod, h36m=h36m, quaternion=quatmod, generator=genmod, source=src)
    _IMPORTED = ns
    return ns


def make_args(batch_size: int, mode: str = "single", architecture: str = "3,3,3", seed: int = 0):
    return argparse.Namespace(batch_size=batch_size, random_seed=seed,
                              single_or_multi_train_mode=mode, architecture=architecture)


# kwarg order of the 15 bone lengths == used_16key_15bone_len_table order
# (forward_kinematics_DH_model.py:46-49,357-361; Fk_generator.py:216-230)
BONE_KWARGS = (
    "left_small_leg_len", "right_small_leg_len", "left_big_leg_len", "right_big_leg_len",
    "left_hip_len", "right_hip_len", "waist_len", "thorax_len", "left_shoulder_len",
    "right_shoulder_len", "left_big_arm_len", "right_big_arm_len", "left_small_arm_len",
    "right_small_arm_len", "neck_len",
)


def ref_fk32(angles33, grot3, bone15, root3):
    """Run the reference torch FK on CPU.  Inputs are torch tensors [N,33],[N,3],[N,15],[N,3].

    Returns the reference's [N,32,3] tensor (autograd graph attached to `angles33`,
    `grot3`, `root3` if they require grad).  A fresh model object is built per call
    because the reference bakes N in at construction and mutates its tables in place.
    """
    ref = import_reference()
    n = angles33.shape[0]
    model = ref.fk.Forward_Kinematics_DH_Model(make_args(n), ["S1"], None)
    kw = dict(
        right_leg_joints_angle=angles33[:, 0:5],
        left_leg_joints_angle=angles33[:, 5:10],
        body_joints_angle=angles33[:, 10:23],
        right_hand_joints_angle=angles33[:, 23:28],
        left_hand_joints_angle=angles33[:, 28:33],
        generator_global_rot_3d_pos_angle=grot3,
        root_3d_pos=root3,
    )
    for i, name in enumerate(BONE_KWARGS):
        kw[name] = bone15[:, i].detach()
    return model.change_3d_joint_angle(**kw)


def ref_pipeline(angles33, grot3, bone15, root3, cam_q, cam_t, cam_intr9):
    """Reference FK -> 32->16 gather -> world->camera -> project (torch CPU, autograd on)."""
    import torch
    ref = import_reference()
    w32 = ref_fk32(angles33, grot3, bone15, root3)
    w16 = w32[:, ref.h36m.H36M_32_To_16_Table]
    n = w16.shape[0]
    q = torch.as_tensor(cam_q, dtype=torch.float32).view(1, 4)
    t = torch.as_tensor(cam_t, dtype=torch.float32).view(1, 3)
    cp = torch.as_tensor(cam_intr9, dtype=torch.float32).view(1, 9).repeat(n, 1)
    cam = ref.camera.GAN_torch_world_to_camera(w16, R=q, t=t)
    uv = ref.camera.project_to_2d(cam, cp)
    return w32, w16, cam, uv


def camera_block(subject: str, cam_id: int):
    """16-float camera block [q(4), t(3) metres, f(2), c(2), k(3), p(2)] built exactly as
    model_fk_gan_train.py:344-363 builds cam_R / cam_t / cam_para_temp (float64 numpy ->
    float32 tensor)."""
    import numpy as np
    ref = import_reference()
    ext = ref.h36m.h36m_cameras_extrinsic_params[subject][cam_id]
    intr = ref.h36m.h36m_cameras_intrinsic_params[cam_id]
    q = np.array(ext["orientation"]).reshape(4)
    t = np.array(ext["translation"]).reshape(3) / 1000.0
    res_w = float(intr["res_w"])
    res_h = float(intr["res_h"])
    f = np.array(intr["focal_length"]) / res_w * 2.0
    c = ref.camera.normalize_screen_coordinates(np.array(intr["center"]), w=res_w, h=res_h).astype("float32")
    k = np.array(intr["radial_distortion"])
    p = np.array(intr["tangential_distortion"])
    blk = np.concatenate([q, t, f, c, k, p]).astype(np.float32)
    assert blk.shape == (16,)
    return blk
