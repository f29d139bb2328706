"""The C-ABI library must load without a GPU and export every symbol include/dhfk.h declares.
No compute is launched here; only argument validation paths (which run before any CUDA call)."""
import ctypes
import os
import re

import numpy as np
import pytest

from dhfk import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "dhfk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dhfk_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    names = _declared_functions()
    assert len(names) >= 12
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libdhfk.so does not export %s" % n
        assert n in _cabi.SIGNATURES, "python binding lacks %s" % n
    assert sorted(_cabi.SIGNATURES) == names


def test_abi_version_and_tile():
    lib = _cabi.load()
    assert lib.dhfk_abi_version() == _cabi.ABI_VERSION == 2
    assert lib.dhfk_tile_rows() == 32


def test_argument_validation_without_gpu():
    lib = _cabi.load()
    z = None
    # n == 0 is a no-op success for every entry point
    assert lib.dhfk_forward(z, 33, z, 3, z, 15, z, 3, z, z, z, z, 0, 0, z) == 0
    assert lib.dhfk_backward(z, 33, z, 3, z, 15, z, 3, z, z, z, z, z, 33, z, 3, z, 3, z, 15, 0, 0, z) == 0
    assert lib.dhfk_world_to_camera_forward(z, z, z, 0, z, 0, z) == 0
    assert lib.dhfk_project_forward(z, z, 9, z, 0, 16, z) == 0
    # negative n, null pointers, short strides -> DHFK_E_INVAL with a message
    assert lib.dhfk_forward(z, 33, z, 3, z, 15, z, 3, z, z, z, z, -1, 0, z) == _cabi.E_INVAL
    assert "n must be" in _cabi.last_error()
    assert lib.dhfk_forward(z, 33, z, 3, z, 15, z, 3, z, z, z, z, 5, 0, z) == _cabi.E_INVAL
    buf = np.zeros(64 * 48, np.float32)
    p = buf.ctypes.data
    assert lib.dhfk_forward(p, 32, p, 3, p, 15, p, 3, z, p, z, z, 4, 0, z) == _cabi.E_INVAL   # ang stride < 33
    assert "stride" in _cabi.last_error()
    assert lib.dhfk_forward(p, 33, p, 3, p, 15, p, 3, z, z, z, z, 4, 0, z) == _cabi.E_INVAL   # out_world missing
    assert lib.dhfk_forward(p, 33, p, 3, p, 15, p, 3, z, p, z, p, 4, 0, z) == _cabi.E_INVAL   # uv without cam
    assert lib.dhfk_forward(p, 33, p, 3, p, 15, p, 3, z, p + 4, z, z, 4, 0, z) == _cabi.E_ALIGN
    both = _cabi.FLAG_FAST_TRIG | _cabi.FLAG_ACCURATE_TRIG          # contradictory trig policies
    assert lib.dhfk_forward(p, 33, p, 3, p, 15, p, 3, z, p, z, z, 4, both, z) == _cabi.E_INVAL
    assert "mutually exclusive" in _cabi.last_error()
    assert lib.dhfk_backward(p, 33, p, 3, p, 15, p, 3, z, z, z, z, p, 33, p, 3, p, 3, z, 15, 4, 0, z) == _cabi.E_INVAL
    assert "upstream" in _cabi.last_error()
    assert lib.dhfk_project_forward(p, p, 8, p, 4, 16, z) == _cabi.E_INVAL
    assert lib.dhfk_host_workspace_bytes(0, 2) == 0
    assert lib.dhfk_host_workspace_bytes(1024, 2) == 1024 * 253 * 4 * 2
    with pytest.raises(ValueError):
        _cabi.check(_cabi.E_INVAL, "x")
    with pytest.raises(NotImplementedError):
        _cabi.check(_cabi.E_UNSUPPORTED, "x")
    with pytest.raises(RuntimeError):
        _cabi.check(700, "x")


def test_argument_validation_of_the_widened_entry_points_without_gpu():
    """SURVEY 8 f2-f4 entry points: every argument error is detected before any CUDA call."""
    lib = _cabi.load()
    z = None
    buf = np.zeros(4096, np.float32)
    p = buf.ctypes.data
    # n == 0: no-op success
    assert lib.dhfk_retarget_project(z, z, z, 5, z, 9, z, z, 0, z) == 0
    assert lib.dhfk_critic_input_forward(z, z, z, 30, 0, 0, z) == 0
    assert lib.dhfk_critic_input_backward(z, z, z, 30, z, 0, 0, z) == 0
    assert lib.dhfk_critic_input_jvp(z, z, z, z, 30, 0, 0, z) == 0
    assert lib.dhfk_flip_pose(z, z, 0, 3, z) == 0
    assert lib.dhfk_bank_gather(z, 96, 9, z, 0, 10, z, z, z, z) == 0
    # retarget
    assert lib.dhfk_retarget_project(p, p, p, 5, p, 9, p, p, -1, z) == _cabi.E_INVAL
    assert lib.dhfk_retarget_project(z, p, p, 5, p, 9, p, p, 4, z) == _cabi.E_INVAL
    assert lib.dhfk_retarget_project(p, p, p, 0, p, 9, p, p, 4, z) == _cabi.E_INVAL          # no templates
    assert lib.dhfk_retarget_project(p, p, p, 5, p, 8, p, p, 4, z) == _cabi.E_INVAL          # cam stride < 9
    assert "9 columns" in _cabi.last_error()
    assert lib.dhfk_retarget_project(p, p, p, 5, z, 9, p, p, 4, z) == _cabi.E_INVAL          # uv without intrinsics
    assert lib.dhfk_retarget_project(p + 4, p, p, 5, p, 9, p, p, 4, z) == _cabi.E_ALIGN
    # critic inputs
    assert lib.dhfk_critic_input_forward(p, p, p, 16, 4, 0, z) == _cabi.E_INVAL              # kcs_cols not 0/15/30
    assert "kcs_cols" in _cabi.last_error()
    assert lib.dhfk_critic_input_forward(p, p, p, 30, 4, 8, z) == _cabi.E_INVAL              # unknown flag
    assert lib.dhfk_critic_input_forward(p, p, z, 30, 4, 0, z) == _cabi.E_INVAL              # kcs wanted, no buffer
    assert lib.dhfk_critic_input_forward(p, z, z, 0, 4, 0, z) == _cabi.E_INVAL               # nothing to compute
    assert lib.dhfk_critic_input_forward(p, p + 4, p, 30, 4, 0, z) == _cabi.E_ALIGN
    assert lib.dhfk_critic_input_backward(p, z, z, 0, p, 4, 0, z) == _cabi.E_INVAL           # no upstream gradient
    assert "upstream" in _cabi.last_error()
    assert lib.dhfk_critic_input_backward(p, p, p, 30, z, 4, 0, z) == _cabi.E_INVAL
    assert lib.dhfk_critic_input_jvp(p, z, p, p, 30, 4, 0, z) == _cabi.E_INVAL               # no tangent
    assert lib.dhfk_critic_input_jvp(p, p, z, z, 0, 4, 0, z) == _cabi.E_INVAL
    # flip
    assert lib.dhfk_flip_pose(p, p + 256, 4, 4, z) == _cabi.E_INVAL                          # dims
    assert lib.dhfk_flip_pose(p + 4, p + 256, 4, 2, z) == _cabi.E_ALIGN
    # bank gather
    assert lib.dhfk_bank_gather(p, 96, 9, z, 4, 10, p, p, p, z) == _cabi.E_INVAL             # no indices
    assert lib.dhfk_bank_gather(p, 88, 9, p, 4, 10, p, p, p, z) == _cabi.E_INVAL             # record too short for 9 cam cols
    assert "rec_floats" in _cabi.last_error()
    assert lib.dhfk_bank_gather(p, 98, 9, p, 4, 10, p, p, p, z) == _cabi.E_INVAL             # not a multiple of 4
    assert lib.dhfk_bank_gather(p, 96, 33, p, 4, 10, p, p, p, z) == _cabi.E_INVAL            # cam_cols > 32
    assert lib.dhfk_bank_gather(p + 4, 96, 9, p, 4, 10, p, p, p, z) == _cabi.E_ALIGN


def test_argument_validation_of_the_gradient_exchange_without_gpu():
    """SURVEY 8 e: dhfk_grad_allreduce rejects bad arguments before any CUDA call."""
    import ctypes
    lib = _cabi.load()
    buf = np.zeros(4096, np.float32)
    p = buf.ctypes.data
    two = (ctypes.c_void_p * 2)(p, p + 1024)
    flags = (ctypes.c_void_p * 2)(p + 2048, p + 4096)
    st = p + 8192
    call = lambda *a: lib.dhfk_grad_allreduce(*a)
    assert call(two, None, flags, st, 0, 2, 0, 0.5, 16, 512, 100, None) == 0                 # empty range: no-op
    assert call(two, None, flags, st, 2, 2, 64, 0.5, 16, 512, 100, None) == _cabi.E_INVAL    # rank outside the world
    assert call(two, None, flags, st, 0, 17, 64, 0.5, 16, 512, 100, None) == _cabi.E_INVAL   # world > DHFK_AR_MAX_WORLD
    assert call(two, None, flags, st, 0, 2, 62, 0.5, 16, 512, 100, None) == _cabi.E_INVAL    # not a multiple of 4
    assert call(two, None, flags, st, 0, 2, 64, 0.5, 0, 512, 100, None) == _cabi.E_INVAL     # max_ctas
    assert call(two, None, flags, st, 0, 2, 64, 0.5, 65, 512, 100, None) == _cabi.E_INVAL
    assert call(two, None, flags, st, 0, 2, 64, 0.5, 16, 500, 100, None) == _cabi.E_INVAL    # cta_threads % 32
    assert call(two, None, flags, st, 0, 2, 64, 0.5, 16, 512, 0, None) == _cabi.E_INVAL      # timeout
    assert call(None, None, flags, st, 0, 2, 64, 0.5, 16, 512, 100, None) == _cabi.E_INVAL
    assert call(two, None, flags, None, 0, 2, 64, 0.5, 16, 512, 100, None) == _cabi.E_INVAL  # no status word
    hole = (ctypes.c_void_p * 2)(p, None)
    assert call(hole, None, flags, st, 0, 2, 64, 0.5, 16, 512, 100, None) == _cabi.E_INVAL
    odd = (ctypes.c_void_p * 2)(p, p + 4)
    assert call(odd, None, flags, st, 0, 2, 64, 0.5, 16, 512, 100, None) == _cabi.E_ALIGN
    assert call(two, p + 4, flags, st, 0, 2, 64, 0.5, 16, 512, 100, None) == _cabi.E_ALIGN   # multicast address
    assert _cabi.AR_FLAG_WORDS == 64 * 2 * 16 + 64
