"""ctypes binding of libdhfk.so (C ABI declared in include/dhfk.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C <pkg>/csrc -j``.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DHFK_LIB_PATH selects an alternative build (A/B measurements of compile-time tunables)
LIB_PATH = os.environ.get("DHFK_LIB_PATH") or os.path.join(_HERE, "lib", "libdhfk.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

ABI_VERSION = 2
FLAG_FAST_TRIG = 0x1
FLAG_ACCURATE_TRIG = 0x2
CRITIC_CENTRE, CRITIC_FLIP = 0x1, 0x2
VIDEO_REVERSE = 0x1
AR_MAX_WORLD, AR_MAX_CTAS = 16, 64
AR_FLAG_WORDS = AR_MAX_CTAS * 2 * AR_MAX_WORLD + AR_MAX_CTAS
E_INVAL, E_ALIGN, E_UNSUPPORTED = -1, -2, -3

_c_f32p = ctypes.c_void_p  # raw device/host addresses are passed as integers
_i64 = ctypes.c_int64
_u32 = ctypes.c_uint32
_vp = ctypes.c_void_p

# name -> (restype, argtypes); must list every symbol include/dhfk.h declares
SIGNATURES = {
    "dhfk_abi_version": (ctypes.c_int, []),
    "dhfk_last_error": (ctypes.c_char_p, []),
    "dhfk_tile_rows": (ctypes.c_int, []),
    "dhfk_topology": (ctypes.c_int, [_vp] * 8),
    "dhfk_forward": (ctypes.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp,
                                    _vp, _vp, _vp, _i64, _u32, _vp]),
    "dhfk_backward": (ctypes.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp,
                                     _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64,
                                     _i64, _u32, _vp]),
    "dhfk_generator_forward": (ctypes.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, ctypes.c_float, _vp,
                                              _vp, _vp, _vp, _i64, _u32, _vp]),
    "dhfk_generator_backward": (ctypes.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, ctypes.c_float, _vp,
                                               _vp, _vp, _vp, _vp, _i64, _i64, _u32, _vp]),
    "dhfk_world_to_camera_forward": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int32, _vp, _i64, _vp]),
    "dhfk_world_to_camera_backward": (ctypes.c_int, [_vp, _vp, ctypes.c_int32, _vp, _i64, _vp]),
    "dhfk_project_forward": (ctypes.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, _vp]),
    "dhfk_project_backward": (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _vp]),
    "dhfk_scatter32_forward": (ctypes.c_int, [_vp, _vp, _i64, _vp, _i64, _vp]),
    "dhfk_scatter32_backward": (ctypes.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "dhfk_retarget_project": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int32, _vp, _i64, _vp, _vp, _i64, _vp]),
    "dhfk_critic_input_forward": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int32, _i64, _u32, _vp]),
    "dhfk_critic_input_backward": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int32, _vp, _i64, _u32, _vp]),
    "dhfk_critic_input_jvp": (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int32, _i64, _u32, _vp]),
    "dhfk_flip_pose": (ctypes.c_int, [_vp, _vp, _i64, ctypes.c_int32, _vp]),
    "dhfk_video_critic_forward": (ctypes.c_int, [_vp, ctypes.c_int32, _u32, _vp, _vp, _vp, _vp, _i64, _vp]),
    "dhfk_video_critic_backward": (ctypes.c_int, [_vp, ctypes.c_int32, _u32, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "dhfk_video_critic_jvp": (ctypes.c_int, [_vp, _vp, ctypes.c_int32, _u32, _vp, _vp, _vp, _vp, _i64, _vp]),
    "dhfk_video_root_diff_forward": (ctypes.c_int, [_vp, ctypes.c_int32, _u32, _vp, _vp, _i64, _vp]),
    "dhfk_video_root_diff_backward": (ctypes.c_int, [_vp, _vp, ctypes.c_int32, _u32, _vp, _i64, _vp]),
    "dhfk_bank_gather": (ctypes.c_int, [_vp, _i64, ctypes.c_int32, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "dhfk_grad_allreduce": (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int32, ctypes.c_int32, _i64, ctypes.c_float,
                                           ctypes.c_int32, ctypes.c_int32, _i64, _vp]),
    "dhfk_host_workspace_bytes": (_i64, [_i64, ctypes.c_int32]),
    "dhfk_forward_backward_host": (ctypes.c_int, [_vp] * 12 + [_i64, _i64, ctypes.c_int32, _vp, _i64, _u32]),
}

_lib = None


def load():
    """Load libdhfk.so (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libdhfk.so not found at %s -- the CUDA extension has not been built and there is no CPU "
            "fallback.  Run `make -C %s -j` (needs nvcc, sm_100a) or __graft_entry__.build()." % (LIB_PATH, CSRC_DIR))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    got = lib.dhfk_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError("libdhfk.so ABI version %d, binding expects %d" % (got, ABI_VERSION))
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().dhfk_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = last_error()
    if rc == E_INVAL:
        raise ValueError("%s: invalid argument: %s" % (what, msg))
    if rc == E_ALIGN:
        raise ValueError("%s: alignment: %s" % (what, msg))
    if rc == E_UNSUPPORTED:
        raise NotImplementedError("%s: %s" % (what, msg))
    raise RuntimeError("%s: CUDA error %d: %s" % (what, rc, msg))
