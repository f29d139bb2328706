"""Drop-in for the critic input transforms of the reference's ``models_Fk_GAN/Fk_discriminator.py``.

``special_KCS_Input_transform`` (Fk_discriminator.py:36-146) and ``video_mode_special_KCS_Input_transform``
(:269-377) keep their names and argument meaning and run as one sm_100a kernel each way (forward, vector-Jacobian
backward, and the Jacobian-vector product that WGAN-GP's ``create_graph=True`` pass differentiates through).  The
critic classes themselves (plain Linear/ReLU stacks on cuBLAS) are the reference's own: they look these two functions
up as module globals, so ``dropin.install(critics=True)`` only rebinds the functions.

``critic_views`` is the opt-in fused form of the train loop's root-centring + flip (model_fk_gan_train.py:311-331).
"""
from __future__ import annotations

from .functional import critic_input, flip_pose  # noqa: F401


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
