"""SURVEY 8 f2, video part: the inputs of the motion critics (per-frame KCS, adjacent-frame differences, playback
reverse).  Goldens: tests/golden/video_critic.npz, produced by the UNMODIFIED reference classes
Video_motion_Fk_3D_Discriminator / Video_motion_Fk_2D_Discriminator (oracle/make_golden.py::video_critic_fixture).

CPU part: the numpy/complex-step oracle (oracle/c_oracle.py::video_critic_*) against the goldens -- features as the
reference's branches receive them, and the VJP against the reference's autograd through the whole critic.
GPU part: the kernels through the C ABI against the goldens and, at ragged sizes / other clip lengths, the oracle."""
import numpy as np
import pytest
import torch

import c_oracle
from conftest import RTOL, assert_parity


@pytest.fixture(scope="module")
def vg(golden):
    return golden("video_critic")


def _args():
    import argparse
    return argparse.Namespace(video_Dis_DenseDim_3D=8, video_Dis_DenseDim_2D=8,
                              motion_Dis_whether_use_3dPos_branch=True, motion_Dis_whether_use_3dDiff_branch=True)


def _critics(vg, device="cpu"):
    from dhfk import Fk_discriminator as fd
    F = int(vg["frames"][0])
    D3 = fd.Video_motion_Fk_3D_Discriminator(device, _args(), F)
    D3.load_state_dict({k[len("d3_w_"):]: torch.tensor(vg[k]) for k in vg if k.startswith("d3_w_")})
    D2 = fd.Video_motion_Fk_2D_Discriminator(device, _args(), F)
    D2.load_state_dict({k[len("d2_w_"):]: torch.tensor(vg[k]) for k in vg if k.startswith("d2_w_")})
    return D3.to(device), D2.to(device), F


def _d3_from_features(D, kcs, dkcs, pos, dpos):
    from dhfk.Fk_discriminator import _stack
    b = kcs.shape[0]
    feats = [_stack(kcs.reshape(b, -1), D.special_KCS_previous, D.special_KCS_block1, D.special_KCS_block2, D.special_KCS_block3),
             _stack(dkcs.reshape(b, -1), D.diff_special_KCS_previous, D.diff_special_KCS_block1, D.diff_special_KCS_block2,
                    D.diff_special_KCS_block3),
             _stack(pos.reshape(b, -1), D.pos_3d_previous, D.pos_3d_block1, D.pos_3d_block2, D.pos_3d_block3),
             _stack(dpos.reshape(b, -1), D.diff_pos_3d_previous, D.diff_pos_3d_block1, D.diff_pos_3d_block2, D.diff_pos_3d_block3)]
    return D.kcs_output(D.kcs_merge_block1(D.kcs_merge_previous(torch.cat(feats, dim=-1))))


def _d2_from_features(D, pos, rdiff):
    from dhfk.Fk_discriminator import _stack
    b = pos.shape[0]
    a = _stack(pos.reshape(b, -1), D.pos_2d_previous, D.pos_2d_block1, D.pos_2d_block2, D.pos_2d_block3)
    c = _stack(rdiff.reshape(b, -1), D.root_diff_2d_previous, D.root_diff_2d_block1, D.root_diff_2d_block2, D.root_diff_2d_block3)
    return D.merge_output(D.merge_block1(D.merge_previous(torch.cat((a, c), dim=-1))))


def _flip_clips(x, F, width):
    return np.ascontiguousarray(np.asarray(x).reshape(-1, F, width)[:, ::-1])


# ------------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("rev", [False, True])
def test_oracle_features_match_reference_branch_inputs(vg, rev):
    F = int(vg["frames"][0])
    tag = "d3_rev" if rev else "d3_fwd"
    o = c_oracle.video_critic_forward(vg["pose"], F, reverse=rev)
    for k in ("kcs", "dkcs", "dpos"):
        assert_parity(o[k], vg["%s_feat_%s" % (tag, k)].reshape(o[k].shape), "%s %s" % (tag, k))
    tag2 = "d2_rev" if rev else "d2_fwd"
    o2 = c_oracle.video_root_diff(vg["uv"], F, reverse=rev)
    assert_parity(o2["rdiff"], vg[tag2 + "_feat_rdiff"].reshape(o2["rdiff"].shape), tag2 + " rdiff")
    # reversal is a permutation of the clip: the position branch receives the flipped clip itself
    assert np.array_equal(o["pos"].astype(np.float32), _flip_clips(vg["pose"], F, 48) if rev else vg["pose"].reshape(-1, F, 48))


@pytest.mark.parametrize("rev", [False, True])
def test_oracle_vjp_matches_reference_autograd_through_the_critic(vg, rev):
    """Gradients of the critic's output w.r.t. its branch inputs come from plain torch layers (seeded state dict);
    pushing them through the oracle's VJP must reproduce the reference's d(out)/d(input clip)."""
    D3, D2, F = _critics(vg)
    tag = "d3_rev" if rev else "d3_fwd"
    # the golden's x is the clip as the critic received it (already flipped for the reverse case); the oracle works on
    # the stored (unflipped) clip with reverse=rev and returns the gradient in storage order
    o = c_oracle.video_critic_forward(vg["pose"], F, reverse=rev)
    feats = {k: torch.tensor(o[k].astype(np.float32), requires_grad=True) for k in ("kcs", "dkcs", "pos", "dpos")}
    d = _d3_from_features(D3, feats["kcs"], feats["dkcs"], feats["pos"], feats["dpos"])
    assert_parity(d.detach().numpy(), vg[tag + "_out"], tag + " critic output", rtol=2e-5)
    gs = torch.autograd.grad((d * torch.tensor(vg[tag + "_g_out"])).sum(), list(feats.values()))
    g = dict(zip(feats.keys(), (x.numpy() for x in gs)))
    gx = c_oracle.video_critic_backward(vg["pose"], F, g_kcs=g["kcs"], g_dkcs=g["dkcs"], g_dpos=g["dpos"], g_pos=g["pos"],
                                        reverse=rev)
    ref = vg[tag + "_g_x"].reshape(-1, F, 48)
    if rev:
        ref = ref[:, ::-1]          # golden gradient is w.r.t. the flipped clip; storage order = flipped back
    scale = max(1.0, float(np.abs(ref).max()))
    assert_parity(gx.reshape(-1, F, 48) / scale, ref / scale, tag + " d(out)/d(clip)", rtol=2e-5)
    # 2-D critic
    tag2 = "d2_rev" if rev else "d2_fwd"
    o2 = c_oracle.video_root_diff(vg["uv"], F, reverse=rev)
    f2 = {k: torch.tensor(o2[k].astype(np.float32), requires_grad=True) for k in ("pos", "rdiff")}
    d2 = _d2_from_features(D2, f2["pos"], f2["rdiff"])
    assert_parity(d2.detach().numpy(), vg[tag2 + "_out"], tag2 + " critic output", rtol=2e-5)
    gp_, gr_ = torch.autograd.grad((d2 * torch.tensor(vg[tag2 + "_g_out"])).sum(), [f2["pos"], f2["rdiff"]])
    gu = c_oracle.video_root_diff_backward(F, g_rdiff=gr_.numpy(), g_pos=gp_.numpy(), reverse=rev)
    ref2 = vg[tag2 + "_g_x"].reshape(-1, F, 32)
    if rev:
        ref2 = ref2[:, ::-1]
    assert_parity(gu.reshape(-1, F, 32), ref2, tag2 + " d(out)/d(clip)", rtol=2e-5)


def test_oracle_jvp_is_adjoint_of_vjp():
    rng = np.random.RandomState(5)
    F, B = 5, 7
    x = (rng.randn(B * F, 16, 3) * 0.4).astype(np.float32)
    v = rng.randn(B * F, 16, 3).astype(np.float32)
    g = dict(g_kcs=rng.randn(B, F, 15).astype(np.float32), g_dkcs=rng.randn(B, F - 1, 15).astype(np.float32),
             g_dpos=rng.randn(B, F - 1, 48).astype(np.float32), g_pos=rng.randn(B, F, 48).astype(np.float32))
    for rev in (False, True):
        t = c_oracle.video_critic_jvp(x, v, F, reverse=rev)
        lhs = (t["kcs"] * g["g_kcs"]).sum() + (t["dkcs"] * g["g_dkcs"]).sum() + (t["dpos"] * g["g_dpos"]).sum() + \
            (t["pos"] * g["g_pos"]).sum()
        rhs = (c_oracle.video_critic_backward(x, F, reverse=rev, **g) * v).sum()
        assert abs(lhs - rhs) <= 1e-9 * max(1.0, abs(lhs))


def test_single_frame_clips_have_empty_differences():
    x = np.random.RandomState(1).randn(6, 16, 3).astype(np.float32)
    o = c_oracle.video_critic_forward(x, 1)
    assert o["dkcs"].shape == (6, 0, 15) and o["dpos"].shape == (6, 0, 48)


def test_c_abi_argument_validation_without_gpu():
    from dhfk import _cabi
    lib = _cabi.load()
    z = None
    assert lib.dhfk_video_critic_forward(z, 9, 0, z, z, z, z, 0, z) == 0
    assert lib.dhfk_video_critic_forward(z, 9, 0, z, z, z, z, 10, z) == _cabi.E_INVAL       # 10 is not a multiple of 9
    assert "multiple of frames" in _cabi.last_error()
    assert lib.dhfk_video_critic_forward(z, 0, 0, z, z, z, z, 9, z) == _cabi.E_INVAL
    assert lib.dhfk_video_critic_forward(z, 9, 2, z, z, z, z, 9, z) == _cabi.E_INVAL        # unknown flag
    buf = np.zeros(9 * 48 + 8, np.float32)
    p = buf.ctypes.data + (-buf.ctypes.data) % 16
    assert lib.dhfk_video_critic_forward(p, 9, 0, p, z, z, z, 9, z) == _cabi.E_INVAL        # dkcs missing, F > 1
    assert lib.dhfk_video_critic_forward(p + 4, 9, 0, p, p, z, z, 9, z) == _cabi.E_ALIGN
    assert lib.dhfk_video_critic_backward(p, 9, 0, z, z, z, z, p, 9, z) == _cabi.E_INVAL     # no upstream gradient
    assert lib.dhfk_video_critic_jvp(p, z, 9, 0, p, p, z, z, 9, z) == _cabi.E_INVAL
    assert lib.dhfk_video_root_diff_forward(z, 9, 0, z, z, 9, z) == _cabi.E_INVAL
    assert lib.dhfk_video_root_diff_backward(z, z, 9, 0, p, 9, z) == _cabi.E_INVAL


# ------------------------------------------------------------------------------------------------------ GPU
def _dev():
    return torch.device("cuda", 0)


@pytest.mark.gpu
@pytest.mark.parametrize("rev", [False, True])
def test_kernel_features_match_reference_branch_inputs(vg, rev):
    import dhfk
    F = int(vg["frames"][0])
    x = torch.tensor(vg["pose"], device=_dev())
    kcs, dkcs, dpos, pos = dhfk.functional.video_critic_input(x, F, reverse=rev, want_dpos=True, want_pos=True)
    tag = "d3_rev" if rev else "d3_fwd"
    assert_parity(kcs.cpu().numpy(), vg[tag + "_feat_kcs"].reshape(kcs.shape), tag + " kcs")
    assert_parity(dkcs.cpu().numpy(), vg[tag + "_feat_dkcs"].reshape(dkcs.shape), tag + " dkcs")
    assert_parity(dpos.cpu().numpy(), vg[tag + "_feat_dpos"].reshape(dpos.shape), tag + " dpos")
    want = _flip_clips(vg["pose"], F, 48) if rev else vg["pose"].reshape(-1, F, 48)
    assert np.array_equal(pos.cpu().numpy(), want)
    u = torch.tensor(vg["uv"], device=_dev())
    rd, pb = dhfk.functional.video_root_diff(u, F, reverse=rev, want_playback=True)
    tag2 = "d2_rev" if rev else "d2_fwd"
    assert_parity(rd.cpu().numpy(), vg[tag2 + "_feat_rdiff"].reshape(rd.shape), tag2 + " rdiff")
    assert np.array_equal(pb.cpu().numpy(), _flip_clips(vg["uv"], F, 32) if rev else vg["uv"].reshape(-1, F, 32))


@pytest.mark.gpu
@pytest.mark.parametrize("rev", [False, True])
def test_motion_critics_reproduce_reference_outputs_gradients_and_penalty(vg, rev):
    """Our classes (reference state dict) on the GPU: critic output, d(out)/d(clip), and the WGAN-GP penalty with all
    its parameter gradients (double backward through the JVP kernel), against the unmodified reference on CPU.
    The reverse case is run both ways: on the flipped clip (what the unmodified train loop feeds) and on the stored
    clip with reverse=True (index math in the kernel)."""
    dev = _dev()
    D3, D2, F = _critics(vg, dev)
    for D, key, width, tagbase in ((D3, "pose", 48, "d3"), (D2, "uv", 32, "d2")):
        tag = tagbase + ("_rev" if rev else "_fwd")
        stored = vg[key].reshape(-1, F, width)
        fed = _flip_clips(stored, F, width) if rev else stored
        g_out = torch.tensor(vg[tag + "_g_out"], device=dev)
        ref_gx = vg[tag + "_g_x"].reshape(-1, F, width)
        scale = max(1.0, float(np.abs(ref_gx).max()))
        runs = [("flipped input", fed, False, ref_gx)]
        if rev:
            runs.append(("reverse flag", stored, True, ref_gx[:, ::-1]))
        for what, xin, flag, want_gx in runs:
            x = torch.tensor(np.ascontiguousarray(xin), device=dev, requires_grad=True)
            d = D(x, reverse=True) if flag else D(x)
            assert_parity(d.detach().cpu().numpy(), vg[tag + "_out"], "%s %s out" % (tag, what), rtol=5e-5)
            (gx,) = torch.autograd.grad((d * g_out).sum(), x)
            assert_parity(gx.cpu().numpy() / scale, np.ascontiguousarray(want_gx) / scale, "%s %s g_x" % (tag, what), rtol=5e-5)
        # WGAN-GP exactly as calc_gradient_penalty (Fk_discriminator.py:208-233) with the golden's alpha
        half = fed.shape[0] // 2
        real, fake = torch.tensor(fed[:half], device=dev).reshape(half, -1), torch.tensor(fed[half:], device=dev).reshape(half, -1)
        alpha = torch.tensor(vg["gp_alpha"], device=dev).expand(real.size())
        inter = (alpha * real + (1 - alpha) * fake).requires_grad_(True)
        D.zero_grad()
        di = D(inter)
        (grads,) = torch.autograd.grad(di, inter, grad_outputs=torch.ones_like(di), create_graph=True, retain_graph=True)
        gp = ((grads.norm(2, dim=1) - 1) ** 2).mean() * 10
        gp.backward()
        assert abs(gp.item() - float(vg[tag + "_gp"][0])) <= 5e-5 * max(1.0, abs(float(vg[tag + "_gp"][0])))
        for k, v in D.named_parameters():
            ref = vg["%s_gpgrad_%s" % (tag, k)]
            got = v.grad.cpu().numpy() if v.grad is not None else np.zeros_like(ref)
            s = max(1.0, float(np.abs(ref).max()))
            assert_parity(got / s, ref / s, "%s gp grad %s" % (tag, k), rtol=5e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("frames,clips", [(1, 37), (2, 50), (9, 1), (9, 57), (27, 13), (3, 4096), (31, 9), (32, 5), (33, 7)])
@pytest.mark.parametrize("rev", [False, True])
def test_kernels_match_oracle_at_ragged_sizes(frames, clips, rev):
    """Tile boundaries (31-row forward tiles, 32-row backward tiles) against every clip length class."""
    from dhfk import _cabi
    lib = _cabi.load()
    dev = _dev()
    rng = np.random.RandomState(frames * 1000 + clips)
    n = frames * clips
    x = (rng.randn(n, 16, 3) * 0.4).astype(np.float32)
    v = rng.randn(n, 16, 3).astype(np.float32)
    g = dict(g_kcs=rng.randn(clips, frames, 15).astype(np.float32), g_dkcs=rng.randn(clips, frames - 1, 15).astype(np.float32),
             g_dpos=rng.randn(clips, frames - 1, 48).astype(np.float32), g_pos=rng.randn(clips, frames, 48).astype(np.float32))
    T = lambda a: torch.tensor(a, device=dev)
    xd, vd = T(x), T(v)
    flags = _cabi.VIDEO_REVERSE if rev else 0
    st = torch.cuda.current_stream().cuda_stream
    P = lambda t: t.data_ptr() if t.numel() else None
    outs = {k: torch.full(s, float("nan"), device=dev) for k, s in (("kcs", (clips, frames, 15)), ("dkcs", (clips, frames - 1, 15)),
                                                                    ("dpos", (clips, frames - 1, 48)), ("pos", (clips, frames, 48)))}
    _cabi.check(lib.dhfk_video_critic_forward(P(xd), frames, flags, P(outs["kcs"]), P(outs["dkcs"]), P(outs["dpos"]),
                                              P(outs["pos"]), n, st), "fwd")
    o = c_oracle.video_critic_forward(x, frames, reverse=rev)
    for k in ("kcs", "dkcs", "dpos", "pos"):
        assert_parity(outs[k].cpu().numpy(), o[k], "forward %s F=%d" % (k, frames))
    touts = {k: torch.full_like(t, float("nan")) for k, t in outs.items()}
    _cabi.check(lib.dhfk_video_critic_jvp(P(xd), P(vd), frames, flags, P(touts["kcs"]), P(touts["dkcs"]), P(touts["dpos"]),
                                          P(touts["pos"]), n, st), "jvp")
    tj = c_oracle.video_critic_jvp(x, v, frames, reverse=rev)
    for k in ("kcs", "dkcs", "dpos", "pos"):
        s = max(1.0, float(np.abs(tj[k]).max())) if tj[k].size else 1.0
        assert_parity(touts[k].cpu().numpy() / s, tj[k] / s, "jvp %s F=%d" % (k, frames))
    gd = {k: T(a) for k, a in g.items()}
    gx = torch.full((n, 16, 3), float("nan"), device=dev)
    _cabi.check(lib.dhfk_video_critic_backward(P(xd), frames, flags, P(gd["g_kcs"]), P(gd["g_dkcs"]), P(gd["g_dpos"]),
                                               P(gd["g_pos"]), P(gx), n, st), "bwd")
    ref = c_oracle.video_critic_backward(x, frames, reverse=rev, **g)
    s = np.maximum(1.0, np.abs(ref).max(axis=(1, 2), keepdims=True))
    assert_parity(gx.cpu().numpy() / s, ref / s, "vjp F=%d" % frames)
    # every subset of upstream gradients the dispatcher knows
    for keys in (("g_kcs",), ("g_kcs", "g_dkcs"), ("g_dpos",), ("g_pos",), ("g_dkcs", "g_dpos", "g_pos")):
        if frames == 1 and not ({"g_kcs", "g_pos"} & set(keys)):
            continue
        sub = {k: (gd[k] if k in keys else None) for k in g}
        gx.fill_(float("nan"))
        _cabi.check(lib.dhfk_video_critic_backward(P(xd), frames, flags, *(P(sub[k]) if sub[k] is not None else None
                                                                         for k in ("g_kcs", "g_dkcs", "g_dpos", "g_pos")),
                                                   P(gx), n, st), "bwd subset")
        ref = c_oracle.video_critic_backward(x, frames, reverse=rev, **{k: (g[k] if k in keys else None) for k in g})
        s = np.maximum(1.0, np.abs(ref).max(axis=(1, 2), keepdims=True))
        assert_parity(gx.cpu().numpy() / s, ref / s, "vjp subset %s F=%d" % ("+".join(keys), frames))
    # 2-D root differences and their transpose
    u = rng.randn(n, 16, 2).astype(np.float32)
    ud = T(u)
    rd = torch.full((clips, frames - 1, 2), float("nan"), device=dev)
    pb = torch.full((clips, frames, 32), float("nan"), device=dev)
    _cabi.check(lib.dhfk_video_root_diff_forward(P(ud), frames, flags, P(rd), P(pb), n, st), "root diff")
    o2 = c_oracle.video_root_diff(u, frames, reverse=rev)
    assert_parity(rd.cpu().numpy(), o2["rdiff"], "root diff F=%d" % frames)
    assert np.array_equal(pb.cpu().numpy(), o2["pos"].astype(np.float32))
    g_rd, g_pb = rng.randn(clips, frames - 1, 2).astype(np.float32), rng.randn(clips, frames, 32).astype(np.float32)
    gu = torch.full((n, 16, 2), float("nan"), device=dev)
    g_rd_d, g_pb_d = T(g_rd), T(g_pb)         # keep them alive: a temporary's block is handed to the next allocation
    _cabi.check(lib.dhfk_video_root_diff_backward(P(g_rd_d), P(g_pb_d), frames, flags, P(gu), n, st), "root diff T")
    torch.cuda.synchronize()
    assert_parity(gu.cpu().numpy(), c_oracle.video_root_diff_backward(frames, g_rd, g_pb, reverse=rev), "root diff T F=%d" % frames)


@pytest.mark.gpu
def test_reverse_flag_equals_flipped_input_through_autograd():
    """reverse=True on the stored clip == the same critic on torch.flip(clip, dims=[1]) (video_GAN_fun.py:222-223)."""
    import dhfk
    dev = _dev()
    F, B = 9, 40
    g = torch.Generator(device=dev).manual_seed(3)
    x = (torch.randn(B, F, 48, generator=g, device=dev) * 0.4).requires_grad_(True)
    outs_r = dhfk.functional.video_critic_input(x, F, reverse=True, want_dpos=True, want_pos=True)
    xf = torch.flip(x, dims=[1]).contiguous()
    outs_f = dhfk.functional.video_critic_input(xf, F, reverse=False, want_dpos=True, want_pos=True)
    ws = [torch.randn(o.shape, generator=g, device=dev) for o in outs_r]
    for a, b in zip(outs_r, outs_f):
        assert torch.equal(a, b)
    (g1,) = torch.autograd.grad(sum((o * w).sum() for o, w in zip(outs_r, ws)), x)
    (g2,) = torch.autograd.grad(sum((o * w).sum() for o, w in zip(outs_f, ws)), x)
    assert_parity(g1.cpu().numpy(), g2.cpu().numpy(), "reverse flag vs flipped input", rtol=RTOL)
