"""Make the unmodified reference (run_Fk_GAN.py and friends) use the native kernels.

The reference has no plugin mechanism -- its modules import each other by name.  ``install()``
therefore (1) replaces ``models_Fk_GAN.forward_kinematics_DH_model.Forward_Kinematics_DH_Model``
with the native class in the already-imported reference module (and in every module that did
``from ... import`` it), and (2) rebinds ``GAN_torch_world_to_camera`` / ``project_to_2d`` in
``common.camera`` and in the caller modules that imported them by name.  Nothing else is touched.
See INTEGRATION.md for the three-line patch on the reference side.
"""
from __future__ import annotations

import sys

from . import camera as _camera
from . import forward_kinematics_DH_model as _fk

_CALLERS = (
    "models_Fk_GAN.model_fk_gan_train", "models_Fk_GAN.video_GAN_fun", "models_Fk_GAN.video_mode_operate",
    "models_Fk_GAN.special_operate", "function_aug.dataloader_update", "run_Fk_GAN", "__main__",
)


def install(verbose: bool = False, generators: bool = False, loader_refresh: bool = False, critics: bool = False):
    """Patch the reference modules that are currently imported.  Idempotent.  Returns the names patched.
    loader_refresh=True also swaps random_bl_aug / video_mode_random_bl_aug / dataloader_update for the fused
    retarget+project versions (same np.random stream, same data_dict contract).
    critics=True rebinds special_KCS_Input_transform / video_mode_special_KCS_Input_transform and gives the reference's
    two motion critics (Video_motion_Fk_3D_Discriminator / Video_motion_Fk_2D_Discriminator, Fk_discriminator.py:381-587)
    the fused `forward` -- same classes, same sub-modules and state dicts, only the method is swapped.
    generators=True also swaps Fk_Generator / Video_Fk_Generator for the fused-epilogue versions (SURVEY 8 f1;
    same constructor and state dict) wherever `my_get_poseFk_model` (model_fk_gan_train.py:97-173) finds them."""
    patched = []
    if critics:   # the critic classes call these two as module globals (Fk_discriminator.py:190, :443)
        from . import Fk_discriminator as _dis
        mod = sys.modules.get("models_Fk_GAN.Fk_discriminator")
        if mod is not None:
            for sym in ("special_KCS_Input_transform", "video_mode_special_KCS_Input_transform"):
                setattr(mod, sym, getattr(_dis, sym))
                patched.append("models_Fk_GAN.Fk_discriminator.%s" % sym)
            for cls, fwd in (("Video_motion_Fk_3D_Discriminator", _dis.video_motion_3d_forward),
                             ("Video_motion_Fk_2D_Discriminator", _dis.video_motion_2d_forward)):
                if hasattr(mod, cls):
                    getattr(mod, cls).forward = fwd
                    patched.append("models_Fk_GAN.Fk_discriminator.%s.forward" % cls)
    if loader_refresh:
        from . import dataloader_update as _du
        for name, syms in (("function_aug.dataloader_update", ("random_bl_aug", "dataloader_update")),
                           ("models_Fk_GAN.video_mode_operate", ("random_bl_aug", "video_mode_random_bl_aug",
                                                                 "video_mode_dataloader_update")),
                           ("run_Fk_GAN", ("dataloader_update", "video_mode_dataloader_update")),
                           ("__main__", ("dataloader_update", "video_mode_dataloader_update"))):
            mod = sys.modules.get(name)
            if mod is None:
                continue
            for sym in syms:
                if hasattr(mod, sym):
                    setattr(mod, sym, getattr(_du, sym))
                    patched.append("%s.%s" % (name, sym))
    if generators:
        from . import Fk_generator as _gen
        for name in ("models_Fk_GAN.Fk_generator", "models_Fk_GAN.model_fk_gan_train", "models_Fk_GAN.video_GAN_fun"):
            mod = sys.modules.get(name)
            if mod is None:
                continue
            for sym in ("Fk_Generator", "Video_Fk_Generator"):
                if hasattr(mod, sym):
                    setattr(mod, sym, getattr(_gen, sym))
                    patched.append("%s.%s" % (name, sym))
    fkmod = sys.modules.get("models_Fk_GAN.forward_kinematics_DH_model")
    if fkmod is not None:
        fkmod.Forward_Kinematics_DH_Model = _fk.Forward_Kinematics_DH_Model
        patched.append("models_Fk_GAN.forward_kinematics_DH_model.Forward_Kinematics_DH_Model")
    cammod = sys.modules.get("common.camera")
    if cammod is not None:
        cammod.GAN_torch_world_to_camera = _camera.GAN_torch_world_to_camera
        cammod.project_to_2d = _camera.project_to_2d
        patched.append("common.camera.{GAN_torch_world_to_camera,project_to_2d}")
    for name in _CALLERS:
        mod = sys.modules.get(name)
        if mod is None:
            continue
        for sym, repl in (("Forward_Kinematics_DH_Model", _fk.Forward_Kinematics_DH_Model),
                          ("GAN_torch_world_to_camera", _camera.GAN_torch_world_to_camera),
                          ("project_to_2d", _camera.project_to_2d)):
            if hasattr(mod, sym):
                setattr(mod, sym, repl)
                patched.append("%s.%s" % (name, sym))
    if verbose:
        for p in patched:
            print("[dhfk] patched", p)
    return patched
