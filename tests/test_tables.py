"""Topology / indexing tables must be BIT-EXACT with the reference (north_star).  Three copies exist
(package tables.py, the constexpr tables compiled into libdhfk.so, the oracle's C tables); each is
compared with tests/golden/tables.npz, which oracle/make_golden.py read off the unmodified reference."""
import ctypes

import numpy as np

from dhfk import _cabi, tables


def _implied_dependency(parent, out16):
    """dep[k, j]: angle j moves output k  <=>  j is a strict ancestor of joint out16[k]
    (a frame's origin does not depend on its own theta)."""
    dep = np.zeros((16, 33), dtype=bool)
    for k, j in enumerate(out16):
        a = parent[j]
        while a >= 0:
            dep[k, a] = True
            a = parent[a]
    return dep


def test_package_tables_bit_exact(golden):
    g = golden("tables")
    assert np.array_equal(tables.ALPHA_DEG, g["alpha"])
    assert np.array_equal(tables.THETA0_DEG, g["theta0"])
    assert np.array_equal(tables.LEN_KIND, g["len_kind"])
    assert np.array_equal(tables.LEN_BONE, g["len_bone"])
    assert np.array_equal(tables.LEN_SIGN, g["len_sign"])
    assert tables.H36M_32_To_16_Table == g["h36m_32_to_16"].tolist()
    assert [tuple(x) for x in g["used_16key_15bone_len_table"].tolist()] == tables.used_16key_15bone_len_table
    assert np.array_equal(tables.GAN_ANGLE_RANGE, g["gan_angle_range"])
    assert np.array_equal(tables.GAN_GLOBAL_ROT_RANGE, g["gan_global_rot_range"])
    assert np.array_equal(tables.BONE_TEMPLATES_GANUTILS_ORDER, g["bone_templates"])
    assert np.array_equal(_implied_dependency(tables.PARENT, tables.OUT16_JOINT), g["dep"])


def test_camera_blocks_bit_exact(golden):
    g = golden("tables")
    for si, s in enumerate(g["camera_subjects"].tolist()):
        for c in range(4):
            blk = tables.camera_block(s, c)
            assert blk.dtype == np.float32
            assert np.array_equal(blk, g["camera_blocks"][si, c]), (s, c)


def test_library_topology_bit_exact(golden):
    g = golden("tables")
    lib = _cabi.load()
    parent = np.empty(33, np.int32); out16 = np.empty(16, np.int32)
    alpha = np.empty(33, np.float32); theta0 = np.empty(33, np.float32)
    kind = np.empty(33, np.int32); bone = np.empty(33, np.int32); sign = np.empty(33, np.int32)
    h = np.empty(16, np.int32)
    rc = lib.dhfk_topology(*(a.ctypes.data for a in (parent, out16, alpha, theta0, kind, bone, sign, h)))
    assert rc == 0
    assert np.array_equal(alpha, g["alpha"]) and np.array_equal(theta0, g["theta0"])
    assert np.array_equal(kind, g["len_kind"]) and np.array_equal(bone, g["len_bone"]) and np.array_equal(sign, g["len_sign"])
    assert np.array_equal(h, g["h36m_32_to_16"])
    assert np.array_equal(parent, tables.PARENT) and np.array_equal(out16, tables.OUT16_JOINT)
    assert np.array_equal(_implied_dependency(parent, out16), g["dep"])
    # NULL pointers are allowed
    assert lib.dhfk_topology(None, None, None, None, None, None, None, None) == 0


def test_oracle_tables_bit_exact(golden, c_oracle):
    g = golden("tables")
    t = c_oracle.tables()
    assert np.array_equal(t["alpha"].astype(np.float32), g["alpha"])
    assert np.array_equal(t["theta0"].astype(np.float32), g["theta0"])
    assert np.array_equal(t["len_kind"], g["len_kind"]) and np.array_equal(t["len_bone"], g["len_bone"])
    assert np.array_equal(t["len_sign"], g["len_sign"]) and np.array_equal(t["h36m_32_to_16"], g["h36m_32_to_16"])
    assert np.array_equal(t["parent"], tables.PARENT) and np.array_equal(t["out16"], tables.OUT16_JOINT)


def test_template_permutation():
    # hip/thigh/shin symmetric pairs land where used_16key_15bone_len_table expects them
    bt = tables.BONE_TEMPLATES
    assert bt.shape == (5, 15)
    assert np.allclose(bt[:, 0], bt[:, 1], atol=1e-4)   # shins
    assert np.allclose(bt[:, 2], bt[:, 3], atol=1e-4)   # thighs
    assert np.allclose(bt[:, 4], bt[:, 5], atol=1e-4)   # hips
    assert (bt[:, 0] > 0.4).all() and (bt[:, 4] < 0.2).all()
