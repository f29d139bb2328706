"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/dhfk_oracle.c (float64 CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  The product package never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libdhfk_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc) if needed; returns the .so path."""
    srcs = [os.path.join(_HERE, f) for f in ("dhfk_oracle.c", "dhfk_oracle_aux.c", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "all"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.dhfk_oracle_num_threads.restype = ctypes.c_int
    return _lib


def _f32(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if cols is not None:
        assert a.ndim == 2 and a.shape[1] == cols, (a.shape, cols)
    return a


def _ptr(a, ty=ctypes.c_float):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ty))


def set_num_threads(n: int) -> None:
    lib().dhfk_oracle_set_num_threads(ctypes.c_int(int(n)))


def num_threads() -> int:
    return int(lib().dhfk_oracle_num_threads())


def forward(ang, grot, bone, root, cam16=None, cam_rows=None, want_world32=False):
    """-> dict(world16 [N,16,3], cam [N,16,3], uv [N,16,2], world32 [N,32,3]) float64 numpy."""
    ang = _f32(ang); grot = _f32(grot, 3); bone = _f32(bone, 15); root = _f32(root, 3)
    n = ang.shape[0]
    assert ang.shape[1] >= 33
    cam16 = None if cam16 is None else _f32(np.asarray(cam16).reshape(16))
    cam_rows = None if cam_rows is None else _f32(cam_rows)
    out = {"world16": np.empty((n, 16, 3), np.float64)}
    if want_world32:
        out["world32"] = np.empty((n, 32, 3), np.float64)
    if cam16 is not None:
        out["cam"] = np.empty((n, 16, 3), np.float64)
        out["uv"] = np.empty((n, 16, 2), np.float64)
    d = ctypes.c_double
    rc = lib().dhfk_oracle_forward(
        ctypes.c_int64(n), _ptr(ang), ctypes.c_int64(ang.shape[1]), _ptr(grot), ctypes.c_int64(3),
        _ptr(bone), ctypes.c_int64(15), _ptr(root), ctypes.c_int64(3), _ptr(cam16), _ptr(cam_rows),
        ctypes.c_int64(0 if cam_rows is None else cam_rows.shape[1]),
        _ptr(out.get("world32"), d), _ptr(out["world16"], d), _ptr(out.get("cam"), d), _ptr(out.get("uv"), d))
    assert rc == 0, rc
    return out


def backward(ang, grot, bone, root, cam16=None, cam_rows=None, g_world=None, g_cam=None, g_uv=None,
             want_bone=True):
    """-> dict(g_ang [N,33], g_grot [N,3], g_root [N,3], g_bone [N,15]) float64 numpy."""
    ang = _f32(ang); grot = _f32(grot, 3); bone = _f32(bone, 15); root = _f32(root, 3)
    n = ang.shape[0]
    cam16 = None if cam16 is None else _f32(np.asarray(cam16).reshape(16))
    cam_rows = None if cam_rows is None else _f32(cam_rows)
    g_world = None if g_world is None else _f32(np.asarray(g_world).reshape(n, 48))
    g_cam = None if g_cam is None else _f32(np.asarray(g_cam).reshape(n, 48))
    g_uv = None if g_uv is None else _f32(np.asarray(g_uv).reshape(n, 32))
    out = {"g_ang": np.empty((n, 33), np.float64), "g_grot": np.empty((n, 3), np.float64),
           "g_root": np.empty((n, 3), np.float64)}
    if want_bone:
        out["g_bone"] = np.empty((n, 15), np.float64)
    d = ctypes.c_double
    rc = lib().dhfk_oracle_backward(
        ctypes.c_int64(n), _ptr(ang), ctypes.c_int64(ang.shape[1]), _ptr(grot), ctypes.c_int64(3),
        _ptr(bone), ctypes.c_int64(15), _ptr(root), ctypes.c_int64(3), _ptr(cam16), _ptr(cam_rows),
        ctypes.c_int64(0 if cam_rows is None else cam_rows.shape[1]),
        _ptr(g_world), _ptr(g_cam), _ptr(g_uv),
        _ptr(out["g_ang"], d), _ptr(out["g_grot"], d), _ptr(out["g_root"], d), _ptr(out.get("g_bone"), d))
    assert rc == 0, rc
    return out


def tables():
    """Restated topology tables in generator joint order (see dhfk_oracle_tables)."""
    alpha = np.empty(33, np.float64); theta0 = np.empty(33, np.float64)
    parent = np.empty(33, np.int32); kind = np.empty(33, np.int32)
    bone = np.empty(33, np.int32); sign = np.empty(33, np.int32)
    out16 = np.empty(16, np.int32); h = np.empty(16, np.int32)
    i32 = ctypes.c_int32
    lib().dhfk_oracle_tables(_ptr(alpha, ctypes.c_double), _ptr(theta0, ctypes.c_double), _ptr(parent, i32),
                             _ptr(kind, i32), _ptr(bone, i32), _ptr(sign, i32), _ptr(out16, i32), _ptr(h, i32))
    return dict(alpha=alpha, theta0=theta0, parent=parent, len_kind=kind, len_bone=bone, len_sign=sign,
                out16=out16, h36m_32_to_16=h)


# ---- generator epilogue (Fk_generator.py:121-230), float64 numpy around the C oracle ----------------------
_GEN_ZERO = (4, 9, 22, 23, 28, 33)


def gen_decode(raw35, half37, mid37, root_scale=10.0):
    """raw [N,35] -> (ang [N,33], grot [N,3], root [N,3], chain) where chain holds d(value)/d(raw column)."""
    raw = np.asarray(raw35, dtype=np.float64)
    n = raw.shape[0]
    t = np.tanh(raw)
    dt = 1.0 - t * t
    slot = np.zeros((n, 37)); dslot = np.zeros((n, 37)); src = -np.ones(37, dtype=np.int64)
    col = 0
    for i in range(37):
        if i in _GEN_ZERO:
            slot[:, i] = mid37[i]
        else:
            slot[:, i] = t[:, col] * half37[i] + mid37[i]
            dslot[:, i] = dt[:, col] * half37[i]
            src[i] = col
            col += 1
    assert col == 31
    root = t[:, 32:35] * root_scale
    droot = dt[:, 32:35] * root_scale
    return slot[:, :33], slot[:, 34:37], root, (src, dslot, droot)


def gen_chain_factor(raw35, half37, mid37, root_scale=10.0):
    """|d(slot or root value)/d(raw column)| per element of the raw output [N,35] (0 for column 31)."""
    _, _, _, (src, dslot, droot) = gen_decode(raw35, half37, mid37, root_scale)
    j = np.zeros((dslot.shape[0], 35))
    for i in range(37):
        if src[i] >= 0:
            j[:, src[i]] = np.abs(dslot[:, i])
    j[:, 32:35] = np.abs(droot)
    return j


def gen_forward(raw35, bone_scaled, half37, mid37, cam16=None, root_scale=10.0):
    ang, grot, root, _ = gen_decode(raw35, half37, mid37, root_scale)
    return forward(ang.astype(np.float32), grot.astype(np.float32), bone_scaled, root.astype(np.float32), cam16)


def gen_backward(raw35, bone_scaled, half37, mid37, cam16=None, g_world=None, g_cam=None, g_uv=None, root_scale=10.0):
    """d/d(raw network output) [N,35] (column 31 = 0)."""
    ang, grot, root, (src, dslot, droot) = gen_decode(raw35, half37, mid37, root_scale)
    b = backward(ang.astype(np.float32), grot.astype(np.float32), bone_scaled, root.astype(np.float32), cam16,
                 g_world=g_world, g_cam=g_cam, g_uv=g_uv, want_bone=False)
    n = ang.shape[0]
    d = np.zeros((n, 35))
    gslot = np.concatenate([b["g_ang"], np.zeros((n, 1)), b["g_grot"]], axis=1)
    for i in range(37):
        if src[i] >= 0:
            d[:, src[i]] = gslot[:, i] * dslot[:, i]
    d[:, 32:35] = b["g_root"] * droot
    return d


def retarget(pose16, templates, tmpl_idx=None, cam_rows=None):
    """random_bl_aug (+ project_to_2d with per-row intrinsics) -> dict(pose [N,16,3], uv [N,16,2]) float64."""
    pose16 = np.ascontiguousarray(pose16, dtype=np.float32).reshape(-1, 16, 3)
    n = pose16.shape[0]
    tm = np.ascontiguousarray(templates, dtype=np.float32).reshape(-1, 15)
    idx = None if tmpl_idx is None else np.ascontiguousarray(tmpl_idx, dtype=np.int32).reshape(n)
    out = {"pose": np.empty((n, 16, 3), np.float64)}
    stride = 0
    if cam_rows is not None:
        cam_rows = _f32(np.asarray(cam_rows).reshape(-1, np.asarray(cam_rows).shape[-1]))
        stride = 0 if cam_rows.shape[0] == 1 and n != 1 else cam_rows.shape[1]
        out["uv"] = np.empty((n, 16, 2), np.float64)
    rc = lib().dhfk_oracle_retarget(
        ctypes.c_int64(n), _ptr(pose16), _ptr(idx, ctypes.c_int32), _ptr(tm), ctypes.c_int32(tm.shape[0]),
        _ptr(cam_rows), ctypes.c_int64(stride), _ptr(out["pose"], ctypes.c_double),
        _ptr(out.get("uv"), ctypes.c_double))
    assert rc == 0, rc
    return out


CRITIC_CENTRE, CRITIC_FLIP = 1, 2


def critic_forward(pose16, flags=0, kcs_cols=30):
    """centre / flip / KCS features -> dict(pos [N,16,3], kcs [N,kcs_cols]) float64."""
    pose16 = np.ascontiguousarray(pose16, dtype=np.float32).reshape(-1, 16, 3)
    n = pose16.shape[0]
    out = {"pos": np.empty((n, 16, 3), np.float64)}
    if kcs_cols:
        out["kcs"] = np.empty((n, kcs_cols), np.float64)
    rc = lib().dhfk_oracle_critic_forward(ctypes.c_int64(n), _ptr(pose16), ctypes.c_uint32(flags),
                                          ctypes.c_int32(kcs_cols), _ptr(out["pos"], ctypes.c_double),
                                          _ptr(out.get("kcs"), ctypes.c_double))
    assert rc == 0, rc
    return out


def critic_jvp(pose16, v, flags=0, kcs_cols=30):
    pose16 = np.ascontiguousarray(pose16, dtype=np.float32).reshape(-1, 16, 3)
    v = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 16, 3)
    n = pose16.shape[0]
    out = {"pos": np.empty((n, 16, 3), np.float64)}
    if kcs_cols:
        out["kcs"] = np.empty((n, kcs_cols), np.float64)
    rc = lib().dhfk_oracle_critic_jvp(ctypes.c_int64(n), _ptr(pose16), _ptr(v), ctypes.c_uint32(flags),
                                      ctypes.c_int32(kcs_cols), _ptr(out["pos"], ctypes.c_double),
                                      _ptr(out.get("kcs"), ctypes.c_double))
    assert rc == 0, rc
    return out


def critic_backward(pose16, g_pos=None, g_kcs=None, flags=0):
    pose16 = np.ascontiguousarray(pose16, dtype=np.float32).reshape(-1, 16, 3)
    n = pose16.shape[0]
    g_pos = None if g_pos is None else np.ascontiguousarray(g_pos, dtype=np.float32).reshape(n, 16, 3)
    kcs_cols = 0
    if g_kcs is not None:
        g_kcs = np.ascontiguousarray(g_kcs, dtype=np.float32).reshape(n, -1)
        kcs_cols = g_kcs.shape[1]
    out = np.empty((n, 16, 3), np.float64)
    rc = lib().dhfk_oracle_critic_backward(ctypes.c_int64(n), _ptr(pose16), _ptr(g_pos), _ptr(g_kcs),
                                           ctypes.c_uint32(flags), ctypes.c_int32(kcs_cols),
                                           _ptr(out, ctypes.c_double))
    assert rc == 0, rc
    return out


def flip(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    dims = x.shape[-1]
    out = np.empty(x.shape, np.float64)
    rc = lib().dhfk_oracle_flip(ctypes.c_int64(x.size // (16 * dims)), _ptr(x), ctypes.c_int32(dims),
                                _ptr(out, ctypes.c_double))
    assert rc == 0, rc
    return out


# ---- SURVEY 8 f2, video part: inputs of the motion critics (numpy float64 on top of the complex-step KCS oracle) ----
def _clips(x, frames, width):
    x = np.asarray(x, dtype=np.float64).reshape(-1, frames, width)
    return x


def _adjacent_diff(x):
    """The reference's loop, restated: out[:, f] = x[:, f+1] - x[:, f] for f in 0..F-2
    (Fk_discriminator.py:457-459, :486-489, :573-576)."""
    b, f, w = x.shape
    out = np.zeros((b, f - 1, w), np.float64)
    for first, second in zip(range(0, f - 1), range(1, f)):
        out[:, first, :] = x[:, second, :] - x[:, first, :]
    return out


def _adjacent_diff_T(g, frames):
    """Transpose of _adjacent_diff: [B,F-1,W] -> [B,F,W]."""
    b, _, w = g.shape
    out = np.zeros((b, frames, w), np.float64)
    for first, second in zip(range(0, frames - 1), range(1, frames)):
        out[:, second, :] += g[:, first, :]
        out[:, first, :] -= g[:, first, :]
    return out


def video_critic_forward(pose16, frames, reverse=False):
    """Video_motion_Fk_3D_Discriminator's branch inputs (Fk_discriminator.py:436-492): dict(kcs [B,F,15],
    dkcs [B,F-1,15], pos [B,F,48], dpos [B,F-1,48]).  reverse: on torch.flip(x, dims=[1]) (video_GAN_fun.py:222-223)."""
    x = np.asarray(pose16, dtype=np.float32).reshape(-1, frames, 48)
    if reverse:
        x = x[:, ::-1]
    x = np.ascontiguousarray(x)
    kcs = critic_forward(x.reshape(-1, 16, 3), 0, 15)["kcs"].reshape(-1, frames, 15)
    xd = x.astype(np.float64)
    return dict(kcs=kcs, dkcs=_adjacent_diff(kcs), pos=xd, dpos=_adjacent_diff(xd))


def video_critic_jvp(pose16, v, frames, reverse=False):
    x = np.asarray(pose16, dtype=np.float32).reshape(-1, frames, 48)
    v = np.asarray(v, dtype=np.float32).reshape(-1, frames, 48)
    if reverse:
        x, v = x[:, ::-1], v[:, ::-1]
    x, v = np.ascontiguousarray(x), np.ascontiguousarray(v)
    tk = critic_jvp(x.reshape(-1, 16, 3), v.reshape(-1, 16, 3), 0, 15)["kcs"].reshape(-1, frames, 15)
    vd = v.astype(np.float64)
    return dict(kcs=tk, dkcs=_adjacent_diff(tk), pos=vd, dpos=_adjacent_diff(vd))


def video_critic_backward(pose16, frames, g_kcs=None, g_dkcs=None, g_dpos=None, g_pos=None, reverse=False):
    """d( <g_kcs,kcs> + <g_dkcs,dkcs> + <g_dpos,dpos> + <g_pos,pos> ) / d pose -> [B*F,16,3] float64 (storage order)."""
    x = np.asarray(pose16, dtype=np.float32).reshape(-1, frames, 48)
    b = x.shape[0]
    if reverse:
        x = x[:, ::-1]
    x = np.ascontiguousarray(x)
    gk = np.zeros((b, frames, 15), np.float64)
    gp = np.zeros((b, frames, 48), np.float64)
    if g_kcs is not None:
        gk += _clips(g_kcs, frames, 15)
    if g_dkcs is not None and frames > 1:
        gk += _adjacent_diff_T(_clips(g_dkcs, frames - 1, 15), frames)
    if g_pos is not None:
        gp += _clips(g_pos, frames, 48)
    if g_dpos is not None and frames > 1:
        gp += _adjacent_diff_T(_clips(g_dpos, frames - 1, 48), frames)
    # the complex-step oracle takes float32 upstream gradients; split the float64 sums into hi + lo parts (linear map)
    def apply(gk_, gp_):
        return critic_backward(x.reshape(-1, 16, 3), g_pos=gp_.reshape(-1, 16, 3), g_kcs=gk_.reshape(-1, 15), flags=0)
    gk_hi, gp_hi = gk.astype(np.float32), gp.astype(np.float32)
    g = apply(gk_hi, gp_hi) + apply((gk - gk_hi).astype(np.float32), (gp - gp_hi).astype(np.float32))
    g = g.reshape(b, frames, 48)
    if reverse:
        g = g[:, ::-1]
    return np.ascontiguousarray(g).reshape(-1, 16, 3)


def video_root_diff(uv16, frames, reverse=False):
    """Video_motion_Fk_2D_Discriminator's branch inputs (Fk_discriminator.py:556-579): dict(pos [B,F,32], rdiff [B,F-1,2])."""
    x = np.asarray(uv16, dtype=np.float64).reshape(-1, frames, 16, 2)
    if reverse:
        x = x[:, ::-1]
    root = np.ascontiguousarray(x[:, :, 0, :])
    return dict(pos=np.ascontiguousarray(x).reshape(-1, frames, 32), rdiff=_adjacent_diff(root))


def video_root_diff_backward(frames, g_rdiff=None, g_pos=None, reverse=False):
    g = None
    if g_pos is not None:
        g = _clips(g_pos, frames, 32).copy()
    if g_rdiff is not None and frames > 1:
        gr = _adjacent_diff_T(_clips(g_rdiff, frames - 1, 2), frames)
        if g is None:
            g = np.zeros((gr.shape[0], frames, 32), np.float64)
        g[:, :, 0:2] += gr
    if reverse:
        g = g[:, ::-1]
    return np.ascontiguousarray(g).reshape(-1, 16, 2)
