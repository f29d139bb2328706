"""Time the kernels either side of the hot path on one GPU: CUDA events, 4 rotating buffer
sets (> L2), algorithmic bytes / time against the measured copy peak.  Prints one JSON object."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dhfk  # noqa: E402
from dhfk import tables  # noqa: E402
from dhfk.functional import _cabi  # noqa: E402


def timeit(fn, sets, steps=50, warmup=5):
    for i in range(warmup):
        fn(sets[i % len(sets)])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        fn(sets[i % len(sets)])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def measure(dev=None, n=1 << 20, peak=None, steps=50):
    """-> dict(poses, peak_gbs, kernels={name: ms, bytes_per_pose, gbs, frac, poses_per_s}).  bench.py imports this for
    its `side_kernels` extra."""
    dev = dev or torch.device("cuda:0")
    if peak is None:
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak = json.load(open(pk)).get("hbm_gbs", 6528.7) if os.path.exists(pk) else 6650.0
    lib = _cabi.load()
    st = torch.cuda.current_stream(dev).cuda_stream
    g = torch.Generator(device=dev).manual_seed(1)
    sets = []
    for _ in range(4):
        pose = torch.randn(n, 16, 3, device=dev, generator=g) * 0.3 + torch.tensor([0.0, 0.0, 4.5], device=dev)
        sets.append(dict(pose=pose, idx=torch.randint(0, 5, (n,), device=dev, dtype=torch.int32, generator=g),
                         cam=torch.tensor(tables.camera_block("S1", 0)[7:16], device=dev).repeat(n, 1).contiguous(),
                         gp=torch.randn(n, 16, 3, device=dev, generator=g), gk=torch.randn(n, 30, device=dev, generator=g),
                         o48=torch.empty(n, 16, 3, device=dev), o32=torch.empty(n, 16, 2, device=dev),
                         o30=torch.empty(n, 30, device=dev)))
    tm = torch.tensor(tables.BONE_TEMPLATES_GANUTILS_ORDER, device=dev)
    P = lambda t: t.data_ptr()
    cases = {
        "retarget+project (f3)": (lambda s: lib.dhfk_retarget_project(P(s["pose"]), P(s["idx"]), P(tm), 5, P(s["cam"]), 9,
                                                                      P(s["o48"]), P(s["o32"]), n, st), 192 + 4 + 36 + 192 + 128),
        "critic fwd centre+flip+kcs30 (f2)": (lambda s: lib.dhfk_critic_input_forward(P(s["pose"]), P(s["o48"]), P(s["o30"]), 30, n, 3, st),
                                              192 + 192 + 120),
        "critic fwd kcs30 only (f2)": (lambda s: lib.dhfk_critic_input_forward(P(s["pose"]), None, P(s["o30"]), 30, n, 0, st), 192 + 120),
        "critic vjp pos+kcs30 (f2)": (lambda s: lib.dhfk_critic_input_backward(P(s["pose"]), P(s["gp"]), P(s["gk"]), 30, P(s["o48"]), n, 1, st),
                                      192 + 192 + 120 + 192),
        "critic jvp pos+kcs30 (f2)": (lambda s: lib.dhfk_critic_input_jvp(P(s["pose"]), P(s["gp"]), P(s["o48"]), P(s["o30"]), 30, n, 1, st),
                                      192 + 192 + 192 + 120),
        "flip 3d (f2)": (lambda s: lib.dhfk_flip_pose(P(s["pose"]), P(s["o48"]), n, 3, st), 384),
        "flip 2d (f2)": (lambda s: lib.dhfk_flip_pose(P(s["gp"]), P(s["o32"]), n, 2, st), 256),
    }
    # the fused forward with all three outputs (the critic step buffers pos_3d_cam as well) and the FK-only forward / backward
    from dhfk import synthetic
    blk = tables.camera_block("S1", 0)
    for i, s_ in enumerate(sets):
        s_.update(synthetic.gan_like_torch(n, dev, seed=50 + i))
        s_["o48b"] = torch.empty(n, 16, 3, device=dev)
        s_["g33"], s_["g3a"], s_["g3b"] = torch.empty(n, 33, device=dev), torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
    cases["FK forward: world + cam + uv"] = (
        lambda s: lib.dhfk_forward(P(s["ang"]), 33, P(s["grot"]), 3, P(s["bone"]), 15, P(s["root"]), 3, blk.ctypes.data,
                                   P(s["o48"]), P(s["o48b"]), P(s["o32"]), n, 0, st), 216 + 192 + 192 + 128)
    cases["FK forward: world only"] = (
        lambda s: lib.dhfk_forward(P(s["ang"]), 33, P(s["grot"]), 3, P(s["bone"]), 15, P(s["root"]), 3, None,
                                   P(s["o48"]), None, None, n, 0, st), 216 + 192)
    cases["FK backward: g_world only"] = (
        lambda s: lib.dhfk_backward(P(s["ang"]), 33, P(s["grot"]), 3, P(s["bone"]), 15, P(s["root"]), 3, None,
                                    P(s["gp"]), None, None, P(s["g33"]), 33, P(s["g3a"]), 3, P(s["g3b"]), 3, None, 15, n, 0, st),
        216 + 192 + 156)
    cases["FK backward: g_world + g_cam + g_uv"] = (
        lambda s: lib.dhfk_backward(P(s["ang"]), 33, P(s["grot"]), 3, P(s["bone"]), 15, P(s["root"]), 3, blk.ctypes.data,
                                    P(s["gp"]), P(s["pose"]), P(s["gu2"]), P(s["g33"]), 33, P(s["g3a"]), 3, P(s["g3b"]), 3, None, 15,
                                    n, 0, st), 216 + 192 + 192 + 128 + 156)
    for s_ in sets:
        s_["gu2"] = torch.randn(n, 16, 2, device=dev, generator=g)
    for s_ in sets:
        s_["g37"] = torch.zeros(n, 37, device=dev)
        s_["g37"][:, :33] = s_["ang"]; s_["g37"][:, 34:] = s_["grot"]
    cases["FK forward world only, angles / global rotation as views of one [N,37] tensor"] = (
        lambda s: lib.dhfk_forward(P(s["g37"]), 37, P(s["g37"]) + 34 * 4, 37, P(s["bone"]), 15, P(s["root"]), 3, None,
                                   P(s["o48"]), None, None, n, 0, st), 148 + 60 + 12 + 192)
    for s_ in sets:
        s_["o96b"] = torch.empty(n, 32, 3, device=dev)
        s_["g32"] = torch.randn(n, 32, 3, device=dev, generator=g)
    cases["32-slot layout forward (change_3d_joint_angle's return tensor)"] = (
        lambda s: lib.dhfk_scatter32_forward(P(s["pose"]), P(s["root"]), 3, P(s["o96b"]), n, st), 192 + 12 + 384)
    cases["32-slot layout backward"] = (
        lambda s: lib.dhfk_scatter32_backward(P(s["g32"]), P(s["o48"]), P(s["g3a"]), n, st), 384 + 192 + 12)
    # standalone camera ops of the drop-in path (common/camera.py:36-38, :62-94)
    blk_dev_q = torch.tensor(tables.camera_block("S1", 0)[0:4], device=dev)
    blk_dev_t = torch.tensor(tables.camera_block("S1", 0)[4:7], device=dev)
    cases["world->camera fwd (standalone)"] = (
        lambda s: lib.dhfk_world_to_camera_forward(P(s["pose"]), P(blk_dev_q), P(blk_dev_t), 1, P(s["o48"]), n * 16, st), 384)
    cases["world->camera bwd (standalone)"] = (
        lambda s: lib.dhfk_world_to_camera_backward(P(s["gp"]), P(blk_dev_q), 1, P(s["o48"]), n * 16, st), 384)
    cases["project fwd, per-row intrinsics (standalone)"] = (
        lambda s: lib.dhfk_project_forward(P(s["pose"]), P(s["cam"]), 9, P(s["o32"]), n, 16, st), 192 + 36 + 128)
    for s_ in sets:
        s_["gu"] = torch.randn(n, 16, 2, device=dev, generator=g)
    cases["project bwd, per-row intrinsics (standalone)"] = (
        lambda s: lib.dhfk_project_backward(P(s["pose"]), P(s["cam"]), 9, P(s["gu"]), P(s["o48"]), n, 16, st), 192 + 36 + 128 + 192)
    # f4: shuffled mini-batch out of a 4x larger device-resident bank (one 384-byte record per pose)
    bank_rows = 4 * n
    rec = torch.randn(bank_rows, 96, device=dev, generator=g)
    for s_ in sets:
        s_["perm"] = torch.randperm(bank_rows, device=dev, generator=g)[:n].contiguous()
        s_["o9"] = torch.empty(n, 9, device=dev)
    cases["bank gather, shuffled batch (f4)"] = (
        lambda s: lib.dhfk_bank_gather(P(rec), 96, 9, P(s["perm"]), n, bank_rows, P(s["o48"]), P(s["o32"]), P(s["o9"]), st),
        8 + 2 * (192 + 128 + 36))
    for s_ in sets:
        s_["o96"] = torch.empty(n, 96, device=dev)
    cases["context: torch.index_select of the same records (384 B in, 384 B out)"] = (
        lambda s: torch.index_select(rec, 0, s["perm"], out=s["o96"]), 8 + 2 * 384)
    # f2, video part: inputs of the motion critics, F = 9 (architecture 3,3), per-frame KCS + adjacent-frame differences
    F = 9
    nv = n // F * F
    bv = nv // F
    for s_ in sets:
        s_["vk"], s_["vdk"] = torch.empty(bv, F, 15, device=dev), torch.empty(bv, F - 1, 15, device=dev)
        s_["vdp"], s_["vgx"] = torch.empty(bv, F - 1, 48, device=dev), torch.empty(nv, 16, 3, device=dev)
        s_["gvk"], s_["gvdk"] = torch.randn(bv, F, 15, device=dev, generator=g), torch.randn(bv, F - 1, 15, device=dev, generator=g)
        s_["gvdp"] = torch.randn(bv, F - 1, 48, device=dev, generator=g)
        s_["vrd"] = torch.empty(bv, F - 1, 2, device=dev)
    fr = (F - 1) / F
    cases["video critic fwd: kcs + dkcs + dpos, F=9 (f2)"] = (
        lambda s: lib.dhfk_video_critic_forward(P(s["pose"]), F, 0, P(s["vk"]), P(s["vdk"]), P(s["vdp"]), None, nv, st),
        192 + 60 + fr * (60 + 192))
    cases["video critic fwd, playback reverse (f2)"] = (
        lambda s: lib.dhfk_video_critic_forward(P(s["pose"]), F, 1, P(s["vk"]), P(s["vdk"]), P(s["vdp"]), None, nv, st),
        192 + 60 + fr * (60 + 192))
    cases["video critic vjp (f2)"] = (
        lambda s: lib.dhfk_video_critic_backward(P(s["pose"]), F, 0, P(s["gvk"]), P(s["gvdk"]), P(s["gvdp"]), None, P(s["vgx"]), nv, st),
        192 + 60 + fr * (60 + 192) + 192)
    cases["video critic jvp (f2)"] = (
        lambda s: lib.dhfk_video_critic_jvp(P(s["pose"]), P(s["gp"]), F, 0, P(s["vk"]), P(s["vdk"]), P(s["vdp"]), None, nv, st),
        192 + 192 + 60 + fr * (60 + 192))
    cases["video 2-D root differences fwd (f2; reads 8 of every 128 bytes: sector-bound)"] = (
        lambda s: lib.dhfk_video_root_diff_forward(P(s["gu"]), F, 0, P(s["vrd"]), None, nv, st), 32 + fr * 8)
    cases["video 2-D root differences bwd (f2)"] = (
        lambda s: lib.dhfk_video_root_diff_backward(P(s["vrd"]), None, F, 0, P(s["o32"]), nv, st), fr * 8 + 128)
    out = {"poses": n, "peak_gbs": peak, "kernels": {}}
    for name, (fn, bpp) in cases.items():
        ms = timeit(fn, sets, steps=steps)
        gbs = n * bpp / ms / 1e6
        out["kernels"][name] = {"ms": round(ms, 5), "bytes_per_pose": bpp, "gbs": round(gbs, 1), "frac": round(gbs / peak, 4),
                                "poses_per_s": round(n / ms * 1e3)}
    return out


if __name__ == "__main__":
    print(json.dumps(measure(n=int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20), indent=1))
