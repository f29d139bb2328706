"""torch.autograd entry points over the C ABI (libdhfk.so).

PyTorch is plumbing here: it owns device memory and streams; every number is produced by the
hand-written sm_100a kernels.  There is no CPU or eager fallback -- CPU tensors are moved to the
CUDA device (as the reference's ``.cuda()`` calls do) and a missing GPU / library raises.

Backward contract: forward saves only its *inputs*; backward recomputes the chain in
registers and is once-differentiable (nothing in the reference differentiates twice through FK:
``calc_gradient_penalty`` works on ``.data``, models_Fk_GAN/model_fk_gan_train.py:213-214).
"""
from __future__ import annotations

import numpy as np
import torch
from torch.autograd.function import once_differentiable

from . import _cabi


def _trig_flags(fast_trig=False, accurate_grad=False) -> int:
    """C-ABI trig flags (include/dhfk.h): default = accurate forward + MUFU backward; fast_trig = MUFU in the forward
    too; accurate_grad = table sincos in the backward too."""
    if fast_trig and accurate_grad:
        raise ValueError("fast_trig and accurate_grad are mutually exclusive")
    return (_cabi.FLAG_FAST_TRIG if fast_trig else 0) | (_cabi.FLAG_ACCURATE_TRIG if accurate_grad else 0)


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("dhfk needs a CUDA device (sm_100a); there is no CPU fallback")


def _stream_ptr(device) -> int:
    """cudaStream_t of torch's current stream on `device` (the raw getter: torch.cuda.current_stream() builds a
    Stream object per call, ~10 us -- comparable to the small-batch kernels themselves)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)


class _on_device:
    """`with torch.cuda.device(d)` only when `d` is not already current (the context manager costs ~5 us)."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        idx = device.index
        self.ctx = None if idx is None or idx == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False


def _rows(t: torch.Tensor, ncols: int, device) -> torch.Tensor:
    """2-D float32 CUDA view [N, >=ncols] with unit column stride (row stride may exceed ncols)."""
    if t.dim() != 2:
        t = t.reshape(-1, t.shape[-1])
    if t.shape[1] < ncols:
        raise ValueError("expected at least %d columns, got %d" % (ncols, t.shape[1]))
    if t.device != device or t.dtype != torch.float32:
        t = t.to(device=device, dtype=torch.float32)
    ok = t.stride(1) == 1 and (t.shape[0] <= 1 or t.stride(0) >= ncols) and t.data_ptr() % 4 == 0
    if not ok:
        t = t.contiguous()
    return t


def _row_stride(t: torch.Tensor) -> int:
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)


def _packed(t, shape, device):
    """Contiguous, 16-byte aligned float32 CUDA tensor of `shape`, or None."""
    if t is None:
        return None
    t = t.to(device=device, dtype=torch.float32).reshape(shape)
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


def cam_block_array(cam) -> np.ndarray:
    """Host float32[16] camera block [q4 (w,x,y,z), t3 (m), f2, c2, k3, p2]."""
    if isinstance(cam, torch.Tensor):
        cam = cam.detach().cpu().numpy()
    a = np.ascontiguousarray(np.asarray(cam, dtype=np.float32).reshape(-1))
    if a.shape[0] != 16:
        raise ValueError("camera block must have 16 floats (q4,t3,f2,c2,k3,p2), got %d" % a.shape[0])
    return a


class _FKProject(torch.autograd.Function):
    """Fused FK (+ camera + projection).  Outputs: world16 [N,16,3] [, cam16 [N,16,3]] [, uv16 [N,16,2]]."""

    @staticmethod
    def forward(ctx, ang, grot, bone, root, cam, want_cam, want_uv, flags):
        _require_cuda()
        lib = _cabi.load()
        device = ang.device if ang.is_cuda else torch.device("cuda", torch.cuda.current_device())
        ang2 = _rows(ang, 33, device)
        n = ang2.shape[0]
        grot2 = _rows(grot, 3, device)
        bone2 = _rows(bone, 15, device)
        root2 = _rows(root, 3, device)
        if not (grot2.shape[0] == n and bone2.shape[0] == n and root2.shape[0] == n):
            raise ValueError("row counts differ: ang %d, grot %d, bone %d, root %d"
                             % (n, grot2.shape[0], bone2.shape[0], root2.shape[0]))
        cam_arr = cam_block_array(cam) if (want_cam or want_uv) else None
        world = torch.empty((n, 16, 3), dtype=torch.float32, device=device)
        camo = torch.empty((n, 16, 3), dtype=torch.float32, device=device) if want_cam else None
        uv = torch.empty((n, 16, 2), dtype=torch.float32, device=device) if want_uv else None
        with _on_device(device):
            rc = lib.dhfk_forward(
                ang2.data_ptr(), _row_stride(ang2), grot2.data_ptr(), _row_stride(grot2),
                bone2.data_ptr(), _row_stride(bone2), root2.data_ptr(), _row_stride(root2),
                cam_arr.ctypes.data if cam_arr is not None else None,
                world.data_ptr(), camo.data_ptr() if want_cam else None, uv.data_ptr() if want_uv else None,
                n, flags, _stream_ptr(device))
        _cabi.check(rc, "dhfk_forward")
        ctx.save_for_backward(ang2, grot2, bone2, root2)
        ctx.cam_arr = cam_arr
        ctx.flags = flags
        ctx.layout = (want_cam, want_uv)
        ctx.in_shapes = (ang.shape, grot.shape, bone.shape, root.shape)
        ctx.in_meta = tuple((t.device, t.dtype) for t in (ang, grot, bone, root))
        outs = (world,) + ((camo,) if want_cam else ()) + ((uv,) if want_uv else ())
        return outs

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        lib = _cabi.load()
        ang2, grot2, bone2, root2 = ctx.saved_tensors
        device = ang2.device
        n = ang2.shape[0]
        want_cam, want_uv = ctx.layout
        it = iter(grads)
        g_world = _packed(next(it), (n, 16, 3), device)
        g_cam = _packed(next(it), (n, 16, 3), device) if want_cam else None
        g_uv = _packed(next(it), (n, 16, 2), device) if want_uv else None
        need_bone = ctx.needs_input_grad[2]
        g_ang = torch.empty((n, 33), dtype=torch.float32, device=device)
        g_grot = torch.empty((n, 3), dtype=torch.float32, device=device)
        g_root = torch.empty((n, 3), dtype=torch.float32, device=device)
        g_bone = torch.empty((n, 15), dtype=torch.float32, device=device) if need_bone else None
        if g_world is None and g_cam is None and g_uv is None:
            g_ang.zero_(); g_grot.zero_(); g_root.zero_()
            if g_bone is not None:
                g_bone.zero_()
        elif n > 0:
            with _on_device(device):
                rc = lib.dhfk_backward(
                    ang2.data_ptr(), _row_stride(ang2), grot2.data_ptr(), _row_stride(grot2),
                    bone2.data_ptr(), _row_stride(bone2), root2.data_ptr(), _row_stride(root2),
                    ctx.cam_arr.ctypes.data if ctx.cam_arr is not None else None,
                    g_world.data_ptr() if g_world is not None else None,
                    g_cam.data_ptr() if g_cam is not None else None,
                    g_uv.data_ptr() if g_uv is not None else None,
                    g_ang.data_ptr(), 33, g_grot.data_ptr(), 3, g_root.data_ptr(), 3,
                    g_bone.data_ptr() if g_bone is not None else None, 15,
                    n, ctx.flags, _stream_ptr(device))
            _cabi.check(rc, "dhfk_backward")

        def back(g, shape, meta, ncols):
            if g is None:
                return None
            if shape[-1] != ncols:  # wider input rows (e.g. a [N,37] angle tensor): pad with zeros
                full = torch.zeros(tuple(shape[:-1]) + (shape[-1],), dtype=torch.float32, device=device)
                full.reshape(-1, shape[-1])[:, :ncols] = g
                g = full
            g = g.reshape(shape)
            if (g.device, g.dtype) != meta:
                g = g.to(device=meta[0], dtype=meta[1])
            return g

        shapes, metas = ctx.in_shapes, ctx.in_meta
        return (back(g_ang if ctx.needs_input_grad[0] else None, shapes[0], metas[0], 33),
                back(g_grot if ctx.needs_input_grad[1] else None, shapes[1], metas[1], 3),
                back(g_bone, shapes[2], metas[2], 15),
                back(g_root if ctx.needs_input_grad[3] else None, shapes[3], metas[3], 3),
                None, None, None, None)


class _FKProjectWide(torch.autograd.Function):
    """The same kernels fed with the generator's own [N,S] tensor (S = 37: angles in columns 0..32, global rotation in
    columns goff..goff+2, Fk_generator.py:136-184) instead of column slices of it: a tile of that tensor is one
    contiguous slab, and the backward returns the gradient of the WHOLE tensor from one slab store -- autograd sees a
    single edge into `wide` instead of six slice views whose backward would each allocate and add a zero [N,S] tensor."""

    @staticmethod
    def forward(ctx, wide, goff, bone, root, cam, want_cam, want_uv, flags):
        _require_cuda()
        lib = _cabi.load()
        device = wide.device
        n, S = wide.shape
        bone2 = _rows(bone, 15, device)
        root2 = _rows(root, 3, device)
        if not (bone2.shape[0] == n and root2.shape[0] == n):
            raise ValueError("row counts differ: angles %d, bone %d, root %d" % (n, bone2.shape[0], root2.shape[0]))
        cam_arr = cam_block_array(cam) if (want_cam or want_uv) else None
        world = torch.empty((n, 16, 3), dtype=torch.float32, device=device)
        camo = torch.empty((n, 16, 3), dtype=torch.float32, device=device) if want_cam else None
        uv = torch.empty((n, 16, 2), dtype=torch.float32, device=device) if want_uv else None
        base = wide.data_ptr()
        if n > 0:
            with _on_device(device):
                rc = lib.dhfk_forward(
                    base, S, base + 4 * goff, S, bone2.data_ptr(), _row_stride(bone2), root2.data_ptr(), _row_stride(root2),
                    cam_arr.ctypes.data if cam_arr is not None else None,
                    world.data_ptr(), camo.data_ptr() if want_cam else None, uv.data_ptr() if want_uv else None,
                    n, flags, _stream_ptr(device))
            _cabi.check(rc, "dhfk_forward")
        ctx.save_for_backward(wide, bone2, root2)
        ctx.cfg = (goff, cam_arr, flags, want_cam, want_uv, bone.shape, root.shape, (bone.device, bone.dtype),
                   (root.device, root.dtype))
        return (world,) + ((camo,) if want_cam else ()) + ((uv,) if want_uv else ())

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        lib = _cabi.load()
        wide, bone2, root2 = ctx.saved_tensors
        goff, cam_arr, flags, want_cam, want_uv, bone_shape, root_shape, bone_meta, root_meta = ctx.cfg
        device = wide.device
        n, S = wide.shape
        it = iter(grads)
        g_world = _packed(next(it), (n, 16, 3), device)
        g_cam = _packed(next(it), (n, 16, 3), device) if want_cam else None
        g_uv = _packed(next(it), (n, 16, 2), device) if want_uv else None
        need_bone = ctx.needs_input_grad[2]
        # full tiles leave as one slab per tile (every column written); a ragged last tile writes only the columns
        # that carry a gradient
        g_wide = (torch.empty if n % 32 == 0 and n > 0 else torch.zeros)((n, S), dtype=torch.float32, device=device)
        g_root = torch.empty((n, 3), dtype=torch.float32, device=device)
        g_bone = torch.empty((n, 15), dtype=torch.float32, device=device) if need_bone else None
        if g_world is None and g_cam is None and g_uv is None:
            g_wide.zero_(); g_root.zero_()
            if g_bone is not None:
                g_bone.zero_()
        elif n > 0:
            base, gbase = wide.data_ptr(), g_wide.data_ptr()
            with _on_device(device):
                rc = lib.dhfk_backward(
                    base, S, base + 4 * goff, S, bone2.data_ptr(), _row_stride(bone2), root2.data_ptr(), _row_stride(root2),
                    cam_arr.ctypes.data if cam_arr is not None else None,
                    g_world.data_ptr() if g_world is not None else None,
                    g_cam.data_ptr() if g_cam is not None else None, g_uv.data_ptr() if g_uv is not None else None,
                    gbase, S, gbase + 4 * goff, S, g_root.data_ptr(), 3,
                    g_bone.data_ptr() if g_bone is not None else None, 15, n, flags, _stream_ptr(device))
            _cabi.check(rc, "dhfk_backward")

        def back(g, shape, meta):
            if g is None:
                return None
            g = g.reshape(shape)
            return g if (g.device, g.dtype) == meta else g.to(device=meta[0], dtype=meta[1])

        return (g_wide if ctx.needs_input_grad[0] else None, None, back(g_bone, bone_shape, bone_meta),
                back(g_root if ctx.needs_input_grad[3] else None, root_shape, root_meta), None, None, None, None)


def wide_rows_ok(wide, goff):
    """True when `wide` [N,S] can be handed to the kernels as it is (see _FKProjectWide)."""
    return (wide.is_cuda and wide.dtype == torch.float32 and wide.dim() == 2 and wide.is_contiguous()
            and 36 <= wide.shape[1] <= 64 and 33 <= goff <= wide.shape[1] - 3 and wide.data_ptr() % 16 == 0)


def fk_world16_wide(wide, goff, bone_len, root, *, fast_trig=False, accurate_grad=False):
    """DH-FK from the generator's [N,S] slot tensor: angles = wide[:, 0:33], global rotation = wide[:, goff:goff+3].
    Returns world16 [N,16,3]; the gradient flows to `wide` as one [N,S] tensor."""
    flags = _trig_flags(fast_trig, accurate_grad)
    return _FKProjectWide.apply(wide, int(goff), bone_len, root, None, False, False, flags)[0]


def fk_project(angles, global_rot, bone_len, root, cam, *, return_cam=True, fast_trig=False, accurate_grad=False):
    """Fused DH-FK -> global rotation/translation -> world->camera -> pinhole projection.

    angles [N,>=33] deg, global_rot [N,3] deg, bone_len [N,15] m, root [N,3] (or [B,F,3]) m,
    cam: 16 floats (see tables.camera_block).  Returns (world16 [N,16,3], cam16 [N,16,3] or None,
    uv16 [N,16,2]).  Replaces Fk_generator.py:232-259 + model_fk_gan_train.py:374-376 in one launch.
    """
    flags = _trig_flags(fast_trig, accurate_grad)
    outs = _FKProject.apply(angles, global_rot, bone_len, root, cam, bool(return_cam), True, flags)
    if return_cam:
        return outs[0], outs[1], outs[2]
    return outs[0], None, outs[1]


def fk_world16(angles, global_rot, bone_len, root, *, fast_trig=False, accurate_grad=False):
    """DH-FK only: world-space 16 joints [N,16,3] (forward_kinematics_DH_model.py:562-822 followed by
    the [:, H36M_32_To_16_Table] gather of Fk_generator.py:259)."""
    flags = _trig_flags(fast_trig, accurate_grad)
    return _FKProject.apply(angles, global_rot, bone_len, root, None, False, False, flags)[0]


class _GeneratorFK(torch.autograd.Function):
    """Generator epilogue + FK (+ camera + projection) in one launch.  Inputs: raw last-layer
    output [N,35], scaled bone lengths [N,15]; outputs like _FKProject.  Gradient flows to the network output only."""

    @staticmethod
    def forward(ctx, net_out, bone, half37, mid37, root_scale, cam, want_cam, want_uv, flags):
        _require_cuda()
        lib = _cabi.load()
        device = net_out.device if net_out.is_cuda else torch.device("cuda", torch.cuda.current_device())
        x = _rows(net_out, 35, device)
        n = x.shape[0]
        bone2 = _rows(bone, 15, device)
        if bone2.shape[0] != n:
            raise ValueError("row counts differ: net_out %d, bone %d" % (n, bone2.shape[0]))
        half = np.ascontiguousarray(np.asarray(half37, dtype=np.float32).reshape(37))
        mid = np.ascontiguousarray(np.asarray(mid37, dtype=np.float32).reshape(37))
        cam_arr = cam_block_array(cam) if (want_cam or want_uv) else None
        world = torch.empty((n, 16, 3), dtype=torch.float32, device=device)
        camo = torch.empty((n, 16, 3), dtype=torch.float32, device=device) if want_cam else None
        uv = torch.empty((n, 16, 2), dtype=torch.float32, device=device) if want_uv else None
        with _on_device(device):
            rc = lib.dhfk_generator_forward(
                x.data_ptr(), _row_stride(x), bone2.data_ptr(), _row_stride(bone2), half.ctypes.data, mid.ctypes.data,
                float(root_scale), cam_arr.ctypes.data if cam_arr is not None else None,
                world.data_ptr(), camo.data_ptr() if want_cam else None, uv.data_ptr() if want_uv else None,
                n, flags, _stream_ptr(device))
        _cabi.check(rc, "dhfk_generator_forward")
        ctx.save_for_backward(x, bone2)
        ctx.consts = (half, mid, float(root_scale), cam_arr, flags, want_cam, want_uv)
        ctx.in_shape, ctx.in_meta = net_out.shape, (net_out.device, net_out.dtype)
        return (world,) + ((camo,) if want_cam else ()) + ((uv,) if want_uv else ())

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        lib = _cabi.load()
        x, bone2 = ctx.saved_tensors
        half, mid, root_scale, cam_arr, flags, want_cam, want_uv = ctx.consts
        device, n = x.device, x.shape[0]
        it = iter(grads)
        g_world = _packed(next(it), (n, 16, 3), device)
        g_cam = _packed(next(it), (n, 16, 3), device) if want_cam else None
        g_uv = _packed(next(it), (n, 16, 2), device) if want_uv else None
        cols = ctx.in_shape[-1]
        g_x = torch.zeros((n, cols), dtype=torch.float32, device=device) if cols != 35 \
            else torch.empty((n, 35), dtype=torch.float32, device=device)
        if g_world is None and g_cam is None and g_uv is None:
            g_x.zero_()
        elif n > 0:
            with _on_device(device):
                rc = lib.dhfk_generator_backward(
                    x.data_ptr(), _row_stride(x), bone2.data_ptr(), _row_stride(bone2), half.ctypes.data,
                    mid.ctypes.data, root_scale, cam_arr.ctypes.data if cam_arr is not None else None,
                    g_world.data_ptr() if g_world is not None else None,
                    g_cam.data_ptr() if g_cam is not None else None, g_uv.data_ptr() if g_uv is not None else None,
                    g_x.data_ptr(), g_x.shape[1], n, flags, _stream_ptr(device))
            _cabi.check(rc, "dhfk_generator_backward")
        g_x = g_x.reshape(ctx.in_shape)
        if (g_x.device, g_x.dtype) != ctx.in_meta:
            g_x = g_x.to(device=ctx.in_meta[0], dtype=ctx.in_meta[1])
        return (g_x,) + (None,) * 8


def generator_fk(net_out, bone_len, *, use_pre_angle=True, root_scale=10.0, cam=None, return_cam=False,
                 return_uv=False, half37=None, mid37=None, fast_trig=False, accurate_grad=False):
    """Fk_generator.py:121-259 after the last Linear layer, fused: net_out [N,35] raw, bone_len [N,15] already
    multiplied by (1 + scaler).  Returns world16 [N,16,3] (and cam16 / uv16 when requested with `cam`)."""
    from . import tables
    if half37 is None or mid37 is None:
        half37, mid37 = tables.generator_slot_scale(use_pre_angle)
    flags = _trig_flags(fast_trig, accurate_grad)
    outs = _GeneratorFK.apply(net_out, bone_len, half37, mid37, root_scale, cam, bool(return_cam), bool(return_uv), flags)
    return outs[0] if len(outs) == 1 else outs


class _WorldToCamera(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, q, t):
        _require_cuda()
        lib = _cabi.load()
        device = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
        xs = x.to(device=device, dtype=torch.float32).contiguous()
        q = q.to(device=device, dtype=torch.float32).reshape(-1).contiguous()
        t = t.to(device=device, dtype=torch.float32).reshape(-1).contiguous()
        out = torch.empty_like(xs)
        npts = xs.numel() // 3
        with _on_device(device):
            rc = lib.dhfk_world_to_camera_forward(xs.data_ptr(), q.data_ptr(), t.data_ptr(), 1, out.data_ptr(),
                                                  npts, _stream_ptr(device))
        _cabi.check(rc, "dhfk_world_to_camera_forward")
        ctx.save_for_backward(q)
        ctx.meta = (x.device, x.dtype)
        if (out.device, out.dtype) != ctx.meta:      # CPU / float64 callers (common/camera.py::wrap) get their own kind back
            out = out.to(device=ctx.meta[0], dtype=ctx.meta[1])
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        lib = _cabi.load()
        (q,) = ctx.saved_tensors
        device = q.device
        g = g.to(device=device, dtype=torch.float32).contiguous()
        gx = torch.empty_like(g)
        with _on_device(device):
            rc = lib.dhfk_world_to_camera_backward(g.data_ptr(), q.data_ptr(), 1, gx.data_ptr(), g.numel() // 3,
                                                   _stream_ptr(device))
        _cabi.check(rc, "dhfk_world_to_camera_backward")
        if (gx.device, gx.dtype) != ctx.meta:
            gx = gx.to(device=ctx.meta[0], dtype=ctx.meta[1])
        return gx, None, None


def world_to_camera(x, q, t):
    """out = qrot(conj(q), x - t) for one camera (q [1,4] or [4], t [1,3] or [3]); x [..., 3].
    Mirrors common/camera.py:36-38 incl. its shape assertions (quaternion.py:17-19)."""
    assert q.shape[-1] == 4
    assert x.shape[-1] == 3
    if q.numel() != 4 or t.numel() != 3:
        raise AssertionError("GAN_torch_world_to_camera expects one camera: R [1,4], t [1,3] "
                             "(the reference's repeat() asserts otherwise, quaternion.py:19)")
    return _WorldToCamera.apply(x, q, t)


class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cam_rows):
        _require_cuda()
        lib = _cabi.load()
        device = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
        xs = x.to(device=device, dtype=torch.float32).contiguous()
        cams = _rows(cam_rows, 9, device)
        n = xs.shape[0]
        joints = xs.numel() // (3 * n) if n > 0 else 0
        uv = torch.empty(tuple(xs.shape[:-1]) + (2,), dtype=torch.float32, device=device)
        with _on_device(device):
            rc = lib.dhfk_project_forward(xs.data_ptr(), cams.data_ptr(), _row_stride(cams), uv.data_ptr(), n, joints,
                                          _stream_ptr(device))
        _cabi.check(rc, "dhfk_project_forward")
        ctx.save_for_backward(xs, cams)
        ctx.meta = (x.device, x.dtype)
        if (uv.device, uv.dtype) != ctx.meta:        # wrap(project_to_2d, True, numpy, numpy) feeds CPU float64 tensors
            uv = uv.to(device=ctx.meta[0], dtype=ctx.meta[1])  # and calls .numpy() on the result (model_fk_gan_train.py:74)
        return uv

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        lib = _cabi.load()
        xs, cams = ctx.saved_tensors
        device = xs.device
        n = xs.shape[0]
        joints = xs.numel() // (3 * n) if n > 0 else 0
        g = g.to(device=device, dtype=torch.float32).contiguous()
        gx = torch.empty_like(xs)
        with _on_device(device):
            rc = lib.dhfk_project_backward(xs.data_ptr(), cams.data_ptr(), _row_stride(cams), g.data_ptr(),
                                           gx.data_ptr(), n, joints, _stream_ptr(device))
        _cabi.check(rc, "dhfk_project_backward")
        if (gx.device, gx.dtype) != ctx.meta:
            gx = gx.to(device=ctx.meta[0], dtype=ctx.meta[1])
        return gx, None


def project_to_2d(x, camera_params):
    """H36M pinhole projection with radial+tangential distortion and the +-1 clamp, per-row
    intrinsics.  Same assertions as common/camera.py:71-74."""
    assert x.shape[-1] == 3
    assert len(camera_params.shape) == 2
    assert camera_params.shape[-1] == 9 or camera_params.shape[-1] == 16
    assert x.shape[0] == camera_params.shape[0]
    return _Project.apply(x, camera_params)


def retarget_project(pose16, templates, tmpl_idx=None, cam_rows=None, *, out_pose=None, out_uv=None):
    """Bone-length retarget (+ per-row projection) in one launch -- SURVEY 8 f3.

    pose16 [N,16,3]; templates [T,15] in utils/gan_utils.py bone order; tmpl_idx [N] int template row per
    pose, or None: every pose takes templates[0] (video_mode_random_bl_aug draws one row per sequence).
    cam_rows [N,>=9] (or a single row [9] / [1,9] shared by all poses) -> also returns uv [N,16,2].
    Replaces random_bl_aug + project_to_2d (function_aug/dataloader_update.py:18-41,69).  Forward only: the
    reference detaches the result before it is used."""
    _require_cuda()
    lib = _cabi.load()
    device = pose16.device if pose16.is_cuda else torch.device("cuda", torch.cuda.current_device())
    assert pose16.shape[-2:] == (16, 3), "pose must be [N,16,3]"
    x = _packed(pose16.detach(), (-1, 16, 3), device)
    n = x.shape[0]
    tm = templates if isinstance(templates, torch.Tensor) else torch.as_tensor(np.asarray(templates, dtype=np.float32))
    tm = tm.detach().to(device=device, dtype=torch.float32).reshape(-1, 15).contiguous()
    idx = None
    if tmpl_idx is not None:
        idx = tmpl_idx if isinstance(tmpl_idx, torch.Tensor) else torch.as_tensor(np.asarray(tmpl_idx))
        idx = idx.to(device=device, dtype=torch.int32).reshape(-1).contiguous()
        if idx.shape[0] != n:
            raise ValueError("tmpl_idx has %d rows, pose has %d" % (idx.shape[0], n))
    cams, cam_stride = None, 0
    if cam_rows is not None:
        cams = cam_rows.detach()
        if cams.dim() == 1 or cams.shape[0] == 1 and n != 1:
            cams = _rows(cams.reshape(1, -1), 9, device)
            cam_stride = 0
        else:
            cams = _rows(cams, 9, device)
            if cams.shape[0] != n:
                raise ValueError("cam_rows has %d rows, pose has %d" % (cams.shape[0], n))
            cam_stride = _row_stride(cams)
    if out_pose is None:
        out_pose = torch.empty((n, 16, 3), dtype=torch.float32, device=device)
    if cams is not None and out_uv is None:
        out_uv = torch.empty((n, 16, 2), dtype=torch.float32, device=device)
    with _on_device(device):
        rc = lib.dhfk_retarget_project(
            x.data_ptr(), idx.data_ptr() if idx is not None else None, tm.data_ptr(), tm.shape[0],
            cams.data_ptr() if cams is not None else None, cam_stride, out_pose.data_ptr(),
            out_uv.data_ptr() if cams is not None else None, n, _stream_ptr(device))
    _cabi.check(rc, "dhfk_retarget_project")
    return (out_pose, out_uv) if cams is not None else out_pose


# ---- critic input transforms ---------------------------------------------------------------------
def _critic_flags(centre, flip):
    return (_cabi.CRITIC_CENTRE if centre else 0) | (_cabi.CRITIC_FLIP if flip else 0)


class _CriticInput(torch.autograd.Function):
    """pose [N,16,3] -> (pos' = centre(flip(pose)), kcs [N,kcs_cols]).  Differentiable twice w.r.t. the upstream
    gradients (what WGAN-GP needs); once w.r.t. the pose."""

    @staticmethod
    def forward(ctx, pose, flags, kcs_cols, want_pos):
        _require_cuda()
        lib = _cabi.load()
        device = pose.device if pose.is_cuda else torch.device("cuda", torch.cuda.current_device())
        x = _packed(pose, (-1, 16, 3), device)
        n = x.shape[0]
        pos = torch.empty((n, 16, 3), dtype=torch.float32, device=device) if want_pos else None
        kcs = torch.empty((n, kcs_cols), dtype=torch.float32, device=device) if kcs_cols else None
        with _on_device(device):
            rc = lib.dhfk_critic_input_forward(x.data_ptr(), pos.data_ptr() if want_pos else None,
                                               kcs.data_ptr() if kcs_cols else None, kcs_cols, n, flags,
                                               _stream_ptr(device))
        _cabi.check(rc, "dhfk_critic_input_forward")
        ctx.save_for_backward(x)
        ctx.cfg = (flags, kcs_cols, want_pos, pose.shape, pose.device, pose.dtype)
        return tuple(o for o in (pos, kcs) if o is not None)

    @staticmethod
    def backward(ctx, *grads):
        (x,) = ctx.saved_tensors
        flags, kcs_cols, want_pos, shape, dev, dtype = ctx.cfg
        it = iter(grads)
        g_pos = next(it) if want_pos else None
        g_kcs = next(it) if kcs_cols else None
        if g_pos is None and g_kcs is None:
            return None, None, None, None
        gx = _CriticInputVJP.apply(x, g_pos, g_kcs, flags, kcs_cols).reshape(shape)
        if (gx.device, gx.dtype) != (dev, dtype):
            gx = gx.to(device=dev, dtype=dtype)
        return gx, None, None, None


class _CriticInputVJP(torch.autograd.Function):
    """g_pose = J(pose)^T (g_pos, g_kcs).  Linear in (g_pos, g_kcs); its derivative w.r.t. them is the JVP kernel.
    The second derivative w.r.t. the pose itself is not propagated (nothing in the reference consumes it: WGAN-GP's
    `interpolates` is a throw-away leaf, Fk_discriminator.py:221-231)."""

    @staticmethod
    def forward(ctx, x, g_pos, g_kcs, flags, kcs_cols):
        lib = _cabi.load()
        device, n = x.device, x.shape[0]
        gp = _packed(g_pos, (n, 16, 3), device)
        gk = _packed(g_kcs, (n, kcs_cols), device) if g_kcs is not None else None
        gx = torch.empty((n, 16, 3), dtype=torch.float32, device=device)
        if n > 0:
            with _on_device(device):
                rc = lib.dhfk_critic_input_backward(x.data_ptr(), gp.data_ptr() if gp is not None else None,
                                                    gk.data_ptr() if gk is not None else None,
                                                    kcs_cols if gk is not None else 0, gx.data_ptr(), n, flags,
                                                    _stream_ptr(device))
            _cabi.check(rc, "dhfk_critic_input_backward")
        ctx.save_for_backward(x)
        ctx.cfg = (flags, kcs_cols, g_pos is not None, g_kcs is not None)
        return gx

    @staticmethod
    @once_differentiable
    def backward(ctx, v):
        lib = _cabi.load()
        (x,) = ctx.saved_tensors
        flags, kcs_cols, has_pos, has_kcs = ctx.cfg
        device, n = x.device, x.shape[0]
        v = _packed(v, (n, 16, 3), device)
        t_pos = torch.empty((n, 16, 3), dtype=torch.float32, device=device) if has_pos else None
        t_kcs = torch.empty((n, kcs_cols), dtype=torch.float32, device=device) if has_kcs else None
        if n > 0:
            with _on_device(device):
                rc = lib.dhfk_critic_input_jvp(x.data_ptr(), v.data_ptr(), t_pos.data_ptr() if has_pos else None,
                                               t_kcs.data_ptr() if has_kcs else None, kcs_cols if has_kcs else 0, n,
                                               flags, _stream_ptr(device))
            _cabi.check(rc, "dhfk_critic_input_jvp")
        return None, t_pos, t_kcs, None, None


def critic_input(pose16, *, centre=False, flip=False, kcs_cols=30, return_pos=True):
    """Fused critic input transform.  pose16 [...,16,3] (or [...,48]) -> (pos' [N,16,3], kcs [N,kcs_cols]):
    pos' = root-centred (model_fk_gan_train.py:312) and/or left-right flipped (:320-327) pose, kcs = the 15 bone-pair
    cosines (+ 15 bone lengths when kcs_cols = 30) of Fk_discriminator.py:36-146 / :269-377 computed on pos'.
    kcs_cols = 0 returns pos' only; return_pos=False returns kcs only.
    Differentiable once w.r.t. the pose and twice w.r.t. the upstream gradients (what WGAN-GP's create_graph=True pass
    needs).  The second derivative w.r.t. the POSE is treated as zero: a gradient penalty differentiated through this
    transform with respect to something upstream of the pose (e.g. generator weights) would be incomplete -- the
    reference never does that (calc_gradient_penalty works on `.data`, model_fk_gan_train.py:213-214)."""
    if kcs_cols not in (0, 15, 30):
        raise ValueError("kcs_cols must be 0, 15 or 30")
    if not return_pos and not kcs_cols:
        raise ValueError("nothing to compute")
    outs = _CriticInput.apply(pose16, _critic_flags(centre, flip), int(kcs_cols), bool(return_pos))
    return outs[0] if len(outs) == 1 else outs


# ---- SURVEY 8 f2, video part: inputs of the motion critics ------------------------------------------------------
def _video_outs(n, frames, device, want_dpos, want_pos):
    b = n // frames
    kcs = torch.empty((b, frames, 15), dtype=torch.float32, device=device)
    dk = torch.empty((b, frames - 1, 15), dtype=torch.float32, device=device)
    dp = torch.empty((b, frames - 1, 48), dtype=torch.float32, device=device) if want_dpos else None
    ps = torch.empty((b, frames, 48), dtype=torch.float32, device=device) if want_pos else None
    return kcs, dk, dp, ps


def _ptr(t):
    return t.data_ptr() if t is not None and t.numel() else None


class _VideoCritic(torch.autograd.Function):
    """pose [B*F,16,3] -> (kcs [B,F,15], dkcs [B,F-1,15][, dpos [B,F-1,48]][, pos [B,F,48]]) in one launch.
    Differentiable twice w.r.t. the upstream gradients (WGAN-GP), once w.r.t. the pose (see _CriticInputVJP)."""

    @staticmethod
    def forward(ctx, pose, frames, flags, want_dpos, want_pos):
        _require_cuda()
        lib = _cabi.load()
        device = pose.device if pose.is_cuda else torch.device("cuda", torch.cuda.current_device())
        x = _packed(pose, (-1, 16, 3), device)
        n = x.shape[0]
        if frames < 1 or n % frames:
            raise ValueError("%d poses are not a whole number of %d-frame clips" % (n, frames))
        kcs, dk, dp, ps = _video_outs(n, frames, device, want_dpos, want_pos)
        if n > 0:
            with _on_device(device):
                rc = lib.dhfk_video_critic_forward(x.data_ptr(), frames, flags, _ptr(kcs), _ptr(dk), _ptr(dp), _ptr(ps), n,
                                                   _stream_ptr(device))
            _cabi.check(rc, "dhfk_video_critic_forward")
        ctx.save_for_backward(x)
        ctx.cfg = (frames, flags, want_dpos, want_pos, pose.shape, pose.device, pose.dtype)
        return tuple(o for o in (kcs, dk, dp, ps) if o is not None)

    @staticmethod
    def backward(ctx, *grads):
        (x,) = ctx.saved_tensors
        frames, flags, want_dpos, want_pos, shape, dev, dtype = ctx.cfg
        it = iter(grads)
        g_kcs, g_dk = next(it), next(it)
        g_dp = next(it) if want_dpos else None
        g_ps = next(it) if want_pos else None
        if g_kcs is None and g_dk is None and g_dp is None and g_ps is None:
            return None, None, None, None, None
        gx = _VideoCriticVJP.apply(x, g_kcs, g_dk, g_dp, g_ps, frames, flags).reshape(shape)
        if (gx.device, gx.dtype) != (dev, dtype):
            gx = gx.to(device=dev, dtype=dtype)
        return gx, None, None, None, None


class _VideoCriticVJP(torch.autograd.Function):
    """g_pose = J(pose)^T (g_kcs, g_dkcs, g_dpos, g_pos): linear in the upstream gradients, so its derivative w.r.t.
    them is the JVP kernel.  The second derivative w.r.t. the pose is not propagated (WGAN-GP's `interpolates` is a
    throw-away leaf, Fk_discriminator.py:221-231): d/d(pose) of a gradient penalty is treated as ZERO.  It cannot be
    refused instead: in the reference's own penalty the pose that reaches this Function is a clone of a view of the
    `interpolates` leaf, so autograd does ask for that gradient there -- only to drop it."""

    @staticmethod
    def forward(ctx, x, g_kcs, g_dk, g_dp, g_ps, frames, flags):
        lib = _cabi.load()
        device, n = x.device, x.shape[0]
        b = n // frames
        gk = _packed(g_kcs, (b, frames, 15), device)
        gdk = _packed(g_dk, (b, frames - 1, 15), device) if frames > 1 else None
        gdp = _packed(g_dp, (b, frames - 1, 48), device) if frames > 1 else None
        gps = _packed(g_ps, (b, frames, 48), device)
        gx = torch.empty((n, 16, 3), dtype=torch.float32, device=device)
        if n > 0:
            with _on_device(device):
                rc = lib.dhfk_video_critic_backward(x.data_ptr(), frames, flags, _ptr(gk), _ptr(gdk), _ptr(gdp), _ptr(gps),
                                                    gx.data_ptr(), n, _stream_ptr(device))
            _cabi.check(rc, "dhfk_video_critic_backward")
        ctx.save_for_backward(x)
        ctx.cfg = (frames, flags, g_kcs is not None, g_dk is not None, g_dp is not None, g_ps is not None)
        return gx

    @staticmethod
    @once_differentiable
    def backward(ctx, v):
        lib = _cabi.load()
        (x,) = ctx.saved_tensors
        frames, flags, has_k, has_dk, has_dp, has_ps = ctx.cfg
        device, n = x.device, x.shape[0]
        v = _packed(v, (n, 16, 3), device)
        t_k, t_dk, t_dp, t_ps = _video_outs(n, frames, device, has_dp, has_ps)
        if n > 0:
            with _on_device(device):
                rc = lib.dhfk_video_critic_jvp(x.data_ptr(), v.data_ptr(), frames, flags, _ptr(t_k), _ptr(t_dk), _ptr(t_dp),
                                               _ptr(t_ps), n, _stream_ptr(device))
            _cabi.check(rc, "dhfk_video_critic_jvp")
        return (None, t_k if has_k else None, t_dk if has_dk else None, t_dp if has_dp else None,
                t_ps if has_ps else None, None, None)


def video_critic_input(pose16, frames, *, reverse=False, want_dpos=True, want_pos=False):
    """Inputs of Video_motion_Fk_3D_Discriminator (Fk_discriminator.py:436-512) from [B*F,16,3] (or [B,F,48]) poses:
    kcs [B,F,15], dkcs [B,F-1,15], then (want_dpos) dpos [B,F-1,48], then (want_pos) pos [B,F,48] -- the clip itself in
    playback order, only worth asking for with reverse=True.  reverse=True yields what the reference computes from
    torch.flip(x.view(B,F,-1), dims=[1]) (video_GAN_fun.py:222-223) without materialising the flipped clip."""
    outs = _VideoCritic.apply(pose16, int(frames), _cabi.VIDEO_REVERSE if reverse else 0, bool(want_dpos), bool(want_pos))
    return outs


class _VideoRootDiff(torch.autograd.Function):
    """uv [B*F,16,2] -> root-joint differences [B,F-1,2] (and the clip in playback order).  Linear: the backward is the
    transpose kernel, and differentiating that (WGAN-GP) is this Function again on the tangent."""

    @staticmethod
    def forward(ctx, uv, frames, flags, want_pb):
        _require_cuda()
        lib = _cabi.load()
        device = uv.device if uv.is_cuda else torch.device("cuda", torch.cuda.current_device())
        x = _packed(uv, (-1, 16, 2), device)
        n = x.shape[0]
        if frames < 1 or n % frames:
            raise ValueError("%d poses are not a whole number of %d-frame clips" % (n, frames))
        b = n // frames
        diff = torch.empty((b, frames - 1, 2), dtype=torch.float32, device=device)
        pb = torch.empty((b, frames, 32), dtype=torch.float32, device=device) if want_pb else None
        if n > 0:
            with _on_device(device):
                rc = lib.dhfk_video_root_diff_forward(x.data_ptr(), frames, flags, _ptr(diff), _ptr(pb), n, _stream_ptr(device))
            _cabi.check(rc, "dhfk_video_root_diff_forward")
        ctx.cfg = (frames, flags, want_pb, n, uv.shape, uv.device, uv.dtype)
        return (diff, pb) if want_pb else (diff,)

    @staticmethod
    def backward(ctx, *grads):
        frames, flags, want_pb, n, shape, dev, dtype = ctx.cfg
        g_diff = grads[0]
        g_pb = grads[1] if want_pb else None
        if g_diff is None and g_pb is None:
            return None, None, None, None
        gx = _VideoRootDiffT.apply(g_diff, g_pb, frames, flags, n).reshape(shape)
        if (gx.device, gx.dtype) != (dev, dtype):
            gx = gx.to(device=dev, dtype=dtype)
        return gx, None, None, None


class _VideoRootDiffT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g_diff, g_pb, frames, flags, n):
        lib = _cabi.load()
        ref = g_diff if g_diff is not None else g_pb
        device = ref.device
        b = n // frames
        gd = _packed(g_diff, (b, frames - 1, 2), device) if frames > 1 else None
        gp = _packed(g_pb, (b, frames, 32), device)
        gx = torch.empty((n, 16, 2), dtype=torch.float32, device=device)
        if gd is None and gp is None:
            gx.zero_()
        elif n > 0:
            with _on_device(device):
                rc = lib.dhfk_video_root_diff_backward(_ptr(gd), _ptr(gp), frames, flags, gx.data_ptr(), n, _stream_ptr(device))
            _cabi.check(rc, "dhfk_video_root_diff_backward")
        ctx.cfg = (frames, flags, g_diff is not None, g_pb is not None)
        return gx

    @staticmethod
    def backward(ctx, v):
        frames, flags, has_d, has_pb = ctx.cfg
        outs = _VideoRootDiff.apply(v, frames, flags, has_pb)
        return (outs[0] if has_d else None, outs[1] if has_pb else None, None, None, None)


def video_root_diff(uv16, frames, *, reverse=False, want_playback=False):
    """Root-joint 2-D differences of Video_motion_Fk_2D_Discriminator (Fk_discriminator.py:566-579): uv16 [B*F,16,2]
    (or [B,F,32]) -> [B,F-1,2]; want_playback=True also returns the clip in playback order [B,F,32]."""
    outs = _VideoRootDiff.apply(uv16, int(frames), _cabi.VIDEO_REVERSE if reverse else 0, bool(want_playback))
    return outs if want_playback else outs[0]


class _FlipPose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _require_cuda()
        lib = _cabi.load()
        device = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
        dims = x.shape[-1]
        xs = x.to(device=device, dtype=torch.float32).contiguous()
        if xs.data_ptr() % 16:
            xs = xs.clone()
        out = torch.empty_like(xs)
        n = xs.numel() // (16 * dims)
        with _on_device(device):
            rc = lib.dhfk_flip_pose(xs.data_ptr(), out.data_ptr(), n, dims, _stream_ptr(device))
        _cabi.check(rc, "dhfk_flip_pose")
        ctx.meta = (x.device, x.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        gx = _FlipPose.apply(g)          # the flip is its own transpose
        if (gx.device, gx.dtype) != ctx.meta:
            gx = gx.to(device=ctx.meta[0], dtype=ctx.meta[1])
        return gx


def flip_pose(x):
    """Left/right flip of [...,16,2|3] keypoints: negate x, swap joints [4,5,6,10,11,12] <-> [1,2,3,13,14,15]
    (model_fk_gan_train.py:320-331, 393-405).  Returns a new tensor (the reference flips a detached clone)."""
    assert x.shape[-2] == 16 and x.shape[-1] in (2, 3), "expected [...,16,2] or [...,16,3]"
    return _FlipPose.apply(x)


class _Scatter32(torch.autograd.Function):
    """world16 [N,16,3], root [N,3] -> the reference's [N,32,3] slot layout, one launch each way."""

    @staticmethod
    def forward(ctx, world16, root):
        _require_cuda()
        lib = _cabi.load()
        device = world16.device if world16.is_cuda else torch.device("cuda", torch.cuda.current_device())
        w = _packed(world16, (-1, 16, 3), device)
        n = w.shape[0]
        r = _rows(root, 3, device)
        if r.shape[0] != n:
            raise ValueError("row counts differ: world16 %d, root %d" % (n, r.shape[0]))
        out = torch.empty((n, 32, 3), dtype=torch.float32, device=device)
        with _on_device(device):
            rc = lib.dhfk_scatter32_forward(w.data_ptr(), r.data_ptr(), _row_stride(r), out.data_ptr(), n, _stream_ptr(device))
        _cabi.check(rc, "dhfk_scatter32_forward")
        ctx.meta = (root.shape, root.device, root.dtype)
        ctx.w_meta = (world16.device, world16.dtype)
        if (out.device, out.dtype) != ctx.w_meta:
            out = out.to(device=ctx.w_meta[0], dtype=ctx.w_meta[1])
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g32):
        lib = _cabi.load()
        device = g32.device if g32.is_cuda else torch.device("cuda", torch.cuda.current_device())
        n = g32.shape[0]
        g = _packed(g32, (n, 32, 3), device)
        g16 = torch.empty((n, 16, 3), dtype=torch.float32, device=device)
        g_root = torch.empty((n, 3), dtype=torch.float32, device=device)
        if n > 0:
            with _on_device(device):
                rc = lib.dhfk_scatter32_backward(g.data_ptr(), g16.data_ptr(), g_root.data_ptr(), n, _stream_ptr(device))
            _cabi.check(rc, "dhfk_scatter32_backward")
        shape, dev, dtype = ctx.meta
        g_root = g_root.reshape(shape)
        if (g_root.device, g_root.dtype) != (dev, dtype):
            g_root = g_root.to(device=dev, dtype=dtype)
        if (g16.device, g16.dtype) != ctx.w_meta:
            g16 = g16.to(device=ctx.w_meta[0], dtype=ctx.w_meta[1])
        return g16, g_root


def scatter_16_to_32(world16, root):
    """[N,16,3] + root [N,3] -> the reference's [N,32,3] layout (forward_kinematics_DH_model.py:745-820): the 16 joints
    in their H36M slots, slot 14 = the head joint again, every other slot = root."""
    return _Scatter32.apply(world16, root)


def fk_project_host(ang, grot, bone, root, cam, g_world=None, g_uv=None, *, chunk_rows=131072, num_streams=3,
                    workspace=None, out=None, fast_trig=False, accurate_grad=False):
    """End-to-end over HOST (ideally pinned) float32 tensors through dhfk_forward_backward_host:
    chunks move through an upload / compute / download stream pipeline over `num_streams` device slots
    (H2D inputs -> fused forward -> D2H world, uv; H2D grads -> fused backward -> D2H grads).
    Returns dict(world, uv[, g_ang, g_grot, g_root]) of host tensors (pinned if allocated here)."""
    _require_cuda()
    lib = _cabi.load()
    n = ang.shape[0]
    for name, t, c in (("ang", ang, 33), ("grot", grot, 3), ("bone", bone, 15), ("root", root, 3)):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != (n, c):
            raise ValueError("%s must be a contiguous float32 host tensor of shape [%d,%d]" % (name, n, c))
    do_bwd = g_world is not None or g_uv is not None
    if do_bwd and (g_world is None or g_uv is None):
        raise ValueError("backward needs both g_world and g_uv")
    if do_bwd:
        for name, t, shape in (("g_world", g_world, (n, 16, 3)), ("g_uv", g_uv, (n, 16, 2))):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != shape:
                raise ValueError("%s must be a contiguous float32 host tensor of shape %r" % (name, shape))
    cam_arr = cam_block_array(cam)
    device = torch.device("cuda", torch.cuda.current_device())
    need = lib.dhfk_host_workspace_bytes(chunk_rows, num_streams)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(need // 4, dtype=torch.float32, device=device)
    if out is None:
        out = {}
    def host(name, shape):
        t = out.get(name)
        if (t is None or tuple(t.shape) != shape or t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()):
            t = torch.empty(shape, dtype=torch.float32, pin_memory=True)
            out[name] = t
        return t
    world = host("world", (n, 16, 3))
    uv = host("uv", (n, 16, 2))
    g_ang = host("g_ang", (n, 33)) if do_bwd else None
    g_grot = host("g_grot", (n, 3)) if do_bwd else None
    g_root = host("g_root", (n, 3)) if do_bwd else None
    p = lambda t: t.data_ptr() if t is not None else None
    with _on_device(device):
        rc = lib.dhfk_forward_backward_host(
            p(ang), p(grot), p(bone), p(root), cam_arr.ctypes.data, p(g_world), p(g_uv), p(world), p(uv),
            p(g_ang), p(g_grot), p(g_root), n, chunk_rows, num_streams, workspace.data_ptr(),
            workspace.numel() * workspace.element_size(), _trig_flags(fast_trig, accurate_grad))
    _cabi.check(rc, "dhfk_forward_backward_host")
    out["_workspace"] = workspace
    return out
