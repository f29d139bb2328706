"""Pin the oracles: the C restatement (float64) and the torch port must reproduce the golden vectors
that oracle/make_golden.py produced by running the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import assert_parity, projection_conditioning

CASES = ["gan133", "stress200", "video36"]


@pytest.mark.parametrize("case", CASES)
def test_c_oracle_forward_matches_reference(golden, c_oracle, case):
    g = golden(case)
    o = c_oracle.forward(g["ang"], g["grot"], g["bone"], g["root"], g["cam_block"], want_world32=True)
    assert_parity(o["world32"], g["world32"], "world32")
    assert_parity(o["world16"], g["world16"], "world16")
    assert_parity(o["cam"], g["cam"], "cam")
    assert_parity(o["uv"], g["uv"], "uv")


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("tag", ["w", "wu", "wcu"])
def test_c_oracle_backward_matches_reference_autograd(golden, c_oracle, case, tag):
    g = golden(case)
    b = c_oracle.backward(g["ang"], g["grot"], g["bone"], g["root"], g["cam_block"], g_world=g["g_world"],
                          g_cam=g["g_cam"] if "c" in tag else None, g_uv=g["g_uv"] if "u" in tag else None)
    cond = projection_conditioning(g["cam"], g["world16"], g["cam_block"], g["g_uv"]) if "u" in tag else None
    assert_parity(b["g_ang"], g["g_ang_" + tag], "g_ang", row_scale=cond)
    assert_parity(b["g_grot"], g["g_grot_" + tag], "g_grot", row_scale=cond)
    assert_parity(b["g_root"], g["g_root_" + tag], "g_root", row_scale=cond)


def test_kat1_tpose_known_answer(golden, c_oracle):
    """init_Fk_DH_angle() of the reference (forward_kinematics_DH_model.py:824-858)."""
    g = golden("kat")
    expect16 = np.array([[0, 0, 0], [.25, 0, 0], [.25, 0, -.6], [.25, 0, -1.1], [-.25, 0, 0], [-.25, 0, -.6],
                         [-.25, 0, -1.1], [0, 0, .25], [0, 0, .45], [0, 0, .6], [-.4, 0, .45], [-.4, 0, .05],
                         [-.4, 0, -.3], [.4, 0, .45], [.4, 0, .05], [.4, 0, -.3]])
    idx = [0, 1, 2, 3, 6, 7, 8, 12, 13, 15, 17, 18, 19, 25, 26, 27]
    assert np.abs(g["tpose32"][idx] - expect16).max() < 1e-6          # the frozen reference output
    bone = np.array([[.5, .5, .6, .6, .25, .25, .25, .2, .4, .4, .4, .4, .35, .35, .15]], np.float32)
    o = c_oracle.forward(np.zeros((1, 33), np.float32), np.zeros((1, 3), np.float32), bone,
                         np.zeros((1, 3), np.float32), want_world32=True)
    assert np.abs(o["world16"][0] - expect16).max() < 1e-7
    assert_parity(o["world32"][0], g["tpose32"], "tpose32")


def test_kat2_bent_pose(golden, c_oracle):
    g = golden("kat")
    for pre in ("bent_", "bent2_"):
        o = c_oracle.forward(g[pre + "ang"], g[pre + "grot"], g[pre + "bone"], g[pre + "root"], g[pre + "cam_block"])
        assert_parity(o["world16"], g[pre + "world16"], pre + "world16")
        assert_parity(o["uv"], g[pre + "uv"], pre + "uv")
    # SURVEY 8c anchors (reference torch fp32 branch)
    w = g["bent_world16"][0]
    assert np.allclose(w[0], [1, 2, 3]) and np.allclose(w[1], [1.1057937, 2.070699, 2.9733663], atol=1e-6)
    assert np.allclose(w[15], [0.9614351, 1.9088607, 3.0104809], atol=1e-6)
    uv = g["bent2_uv"][0]
    assert np.allclose(uv[0], [-0.1133665, -0.1590322], atol=1e-6)
    assert np.allclose(uv[3], [0.0950191, 0.1502595], atol=1e-6)


def test_kat3_structure(golden, c_oracle):
    g = golden("gan133")
    for tag in ("w", "wu", "wcu"):
        assert np.all(g["g_ang_" + tag][:, [4, 9, 22, 27, 32]] == 0)      # chain ends never move an output
    w32 = g["world32"]
    assert np.array_equal(w32[:, 14], w32[:, 15])                        # slot 14 duplicates the head
    unused = [s for s in range(32) if s not in (0, 1, 2, 3, 6, 7, 8, 12, 13, 14, 15, 17, 18, 19, 25, 26, 27)]
    assert np.array_equal(w32[:, unused], np.broadcast_to(g["root"][:, None, :], (133, len(unused), 3)))
    b = c_oracle.backward(g["ang"], g["grot"], g["bone"], g["root"], g["cam_block"], g_world=g["g_world"], g_uv=g["g_uv"])
    assert np.all(b["g_ang"][:, [4, 9, 22, 27, 32]] == 0)


def test_oracle_bone_gradient_finite_difference(golden, c_oracle):
    """The reference never differentiates bone lengths; check the oracle's d/d(bone) against central
    differences of its own forward (float64), so the kernels' optional g_bone has a checker."""
    g = golden("gan133")
    n = 8
    sl = slice(0, n)
    args = (g["ang"][sl], g["grot"][sl])
    gw, gu = g["g_world"][sl].astype(np.float64), g["g_uv"][sl].astype(np.float64)
    b = c_oracle.backward(*args, g["bone"][sl], g["root"][sl], g["cam_block"], g_world=gw, g_uv=gu)
    eps = 2.0 ** -10   # exactly representable perturbation
    for j in range(15):
        bp = g["bone"][sl].copy(); bm = g["bone"][sl].copy()
        bp[:, j] += eps; bm[:, j] -= eps
        op = c_oracle.forward(*args, bp, g["root"][sl], g["cam_block"])
        om = c_oracle.forward(*args, bm, g["root"][sl], g["cam_block"])
        dp = (bp[:, j].astype(np.float64) - bm[:, j].astype(np.float64))
        fd = (((op["world16"] - om["world16"]) * gw).sum((1, 2)) + ((op["uv"] - om["uv"]) * gu).sum((1, 2))) / dp
        assert np.abs(fd - b["g_bone"][:, j]).max() < 5e-4 * max(1.0, np.abs(fd).max()), j


@pytest.mark.parametrize("case", CASES)
def test_torch_port_matches_reference(golden, case):
    """The torch port issues the reference's op sequence; on the generating host it is bit-identical,
    elsewhere libm / BLAS rounding may differ in the last ulp."""
    import torch_port as tp
    g = golden(case)
    torch.set_num_threads(2)
    ang = torch.tensor(g["ang"], requires_grad=True); grot = torch.tensor(g["grot"], requires_grad=True)
    root = torch.tensor(g["root"], requires_grad=True); bone = torch.tensor(g["bone"])
    w32, w16, cam, uv = tp.pipeline(ang, grot, bone, root, g["cam_block"])
    for name, x in (("world32", w32), ("cam", cam), ("uv", uv)):
        assert_parity(x.detach().numpy(), g[name], name, rtol=2e-6)
    loss = (w16 * torch.tensor(g["g_world"])).sum() + (cam * torch.tensor(g["g_cam"])).sum() + (uv * torch.tensor(g["g_uv"])).sum()
    loss.backward()
    cond = projection_conditioning(g["cam"], g["world16"], g["cam_block"], g["g_uv"])
    assert_parity(ang.grad.numpy(), g["g_ang_wcu"], "g_ang", rtol=2e-6, row_scale=cond)
    assert_parity(grot.grad.numpy(), g["g_grot_wcu"], "g_grot", rtol=2e-6, row_scale=cond)
    assert_parity(root.grad.numpy(), g["g_root_wcu"], "g_root", rtol=2e-6, row_scale=cond)


def test_oracle_camera_ops(golden, c_oracle):
    """project_to_2d with per-row intrinsics incl. clamped points (common/camera.py:62-94)."""
    g = golden("camera_ops")
    assert g["clamped"].any()
    assert np.array_equal(g["uv"], g["uv16"])     # the reference only reads the first 9 of 16 columns


def _scaled_bone(bone, scaler):
    from dhfk import tables
    grp = tables.BONE_SCALER_GROUP
    return (bone * np.where(grp[None, :] >= 0, 1.0 + scaler[:, np.maximum(grp, 0)], 1.0)).astype(np.float32)


@pytest.mark.parametrize("tag,pre", [("single", True), ("single_nopre", False), ("video", True)])
def test_generator_epilogue_oracle_matches_reference(golden, c_oracle, tag, pre):
    """the generator epilogue restated in the oracle (tanh, 31->37 slot scatter, range map, x10 root,
    bone scaler) reproduces the reference's Fk_Generator / Video_Fk_Generator outputs and d/d(raw output)."""
    from dhfk import tables
    g = golden("generator")
    half, mid = tables.generator_slot_scale(pre)
    raw = g[tag + "_raw"].reshape(-1, 35)
    bone = _scaled_bone(g[tag + "_bone"], g[tag + "_scaler"])
    o = c_oracle.gen_forward(raw, bone, half, mid)
    fake = g[tag + "_fake"].reshape(-1, 16, 3)
    assert_parity(o["world16"], fake, "fake")
    d = c_oracle.gen_backward(raw, bone, half, mid, g_world=g[tag + "_g_fake"].reshape(-1, 16, 3))
    ref = g[tag + "_d_raw"].reshape(-1, 35)
    assert_parity(d, ref, "d_raw")
    assert np.all(ref[:, 31] == 0)                      # network column 31 is never consumed
    assert np.all(ref[:, [23, 27]] == 0)                # columns feeding the chain-end slots 27 / 32
    assert np.array_equal(tables.generator_src_col()[[27, 32, 34, 36]], [23, 27, 28, 30])


def test_torch_port_kcs_matches_reference(golden):
    """oracle/torch_port.py::special_kcs (the eager context baseline of tools/gan_step_bench.py) against the golden
    produced by the reference's special_KCS_Input_transform, outputs and gradients."""
    import torch_port
    g = golden("critic")
    x = torch.tensor(g["c1f0_pos"], requires_grad=True)
    k = torch_port.special_kcs(x)
    assert_parity(k.detach().numpy(), g["c1f0_kcs"], "kcs")
    x0 = torch.tensor(g["pose"], requires_grad=True)
    (torch_port.special_kcs(x0) * torch.tensor(g["g_kcs"])).sum().backward()
    assert_parity(x0.grad.numpy(), g["c0f0_g_pose_kcs_only"], "g_pose")
