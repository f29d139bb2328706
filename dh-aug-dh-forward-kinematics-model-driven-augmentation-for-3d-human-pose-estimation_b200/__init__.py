"""dhfk -- B200-native DH forward kinematics + camera projection (forward & backward) for DH-AUG.

Public surface:
  fk_project, fk_world16, generator_fk, world_to_camera, project_to_2d, fk_project_host   (functional.py)
  Fk_generator.Fk_Generator / Video_Fk_Generator                            (reference-shaped generators, fused epilogue)
  Forward_Kinematics_DH_Model                                               (reference-shaped class)
  camera.GAN_torch_world_to_camera / camera.project_to_2d                   (reference-shaped functions)
  dataloader_update.random_bl_aug / dataloader_update / refresh_poses        (per-epoch loader refresh)
  Fk_discriminator.special_KCS_Input_transform / critic_views, functional.critic_input / flip_pose   
  pose_buffer.DevicePoseBuffer / shuffled_order / gather_pairs              (device-resident fake-pair bank)
  dropin.install()                                                          (patch the imported reference)
  tables, synthetic, parallel

Importing this package does not touch the GPU and does not load libdhfk.so; the first call does,
and raises if the library is missing (no CPU fallback).
"""
from . import _cabi, tables  # noqa: F401


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `import dhfk` stays cheap
    import importlib
    def keep(value):                 # resolve once: later lookups hit the module dict, not this hook
        globals()[name] = value
        return value
    if name in ("functional", "camera", "dropin", "parallel", "synthetic", "forward_kinematics_DH_model",
                "Fk_generator", "dataloader_update", "Fk_discriminator",
                "pose_buffer"):
        return keep(importlib.import_module("." + name, __name__))
    if name in ("fk_project", "fk_world16", "world_to_camera", "project_to_2d", "fk_project_host", "generator_fk",
                "retarget_project", "critic_input", "flip_pose"):
        return keep(getattr(importlib.import_module(".functional", __name__), name))
    if name == "Forward_Kinematics_DH_Model":
        return keep(importlib.import_module(".forward_kinematics_DH_model", __name__).Forward_Kinematics_DH_Model)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
