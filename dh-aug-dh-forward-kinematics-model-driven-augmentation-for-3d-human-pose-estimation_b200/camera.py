"""Drop-in for the hot functions of the reference's ``common/camera.py``.

``GAN_torch_world_to_camera`` (common/camera.py:36-38) and ``project_to_2d`` (:62-94) keep their
names, argument meaning and assertion behaviour, and run as sm_100a kernels with analytic
backward.  ``install()`` in dropin.py patches them into the reference's own module, leaving every
other helper of that module untouched.
"""
from __future__ import annotations

from .functional import project_to_2d as _project_to_2d
from .functional import world_to_camera as _world_to_camera


def GAN_torch_world_to_camera(X, R, t):
    """X [N,16,3] world-space, R [1,4] camera orientation quaternion (w,x,y,z), t [1,3] metres
    -> camera-space [N,16,3] = qrot(qinverse(R), X - t)."""
    return _world_to_camera(X, R, t)


def project_to_2d(X, camera_params):
    """X [N,*,3] camera-space, camera_params [N,9|16] = f2,c2,k3,p2 -> [N,*,2] normalised screen coords."""
    return _project_to_2d(X, camera_params)
