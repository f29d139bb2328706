#!/usr/bin/env python
"""Launch ONE kernel family a few times so that `ncu --set full -k regex:... -c N` captures it in isolation.
usage: python tools/ncu_target.py {gen|bank|video|wide} [n]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dhfk  # noqa: E402
from dhfk import _cabi, synthetic, tables  # noqa: E402

which = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
dev = torch.device("cuda", 0)
lib = _cabi.load()
st = torch.cuda.current_stream().cuda_stream
blk = tables.camera_block("S1", 0)
P = lambda t: t.data_ptr()
g = torch.Generator(device=dev).manual_seed(1)
d = synthetic.gan_like_torch(n, dev, seed=2)
gw = torch.randn((n, 16, 3), generator=g, device=dev)
gu = torch.randn((n, 16, 2), generator=g, device=dev)
world = torch.empty((n, 16, 3), device=dev)
uv = torch.empty((n, 16, 2), device=dev)
reps = 4
if which == "gen":
    half, mid = tables.generator_slot_scale(True)
    raw = torch.randn((n, 35), generator=g, device=dev)
    raw[:, 32:35] = torch.rand((n, 3), generator=g, device=dev) * 0.2 - 0.1
    raw[:, 34] += 0.1
    d_raw = torch.empty((n, 35), device=dev)
    for _ in range(reps):
        _cabi.check(lib.dhfk_generator_forward(P(raw), 35, P(d["bone"]), 15, half.ctypes.data, mid.ctypes.data, 10.0,
                                               blk.ctypes.data, P(world), None, P(uv), n, 0, st), "gen fwd")
        _cabi.check(lib.dhfk_generator_backward(P(raw), 35, P(d["bone"]), 15, half.ctypes.data, mid.ctypes.data, 10.0,
                                                blk.ctypes.data, P(gw), None, P(gu), P(d_raw), 35, n, 0, st), "gen bwd")
elif which == "bank":
    from dhfk import pose_buffer
    rows = 4 * n
    bank = torch.randn((rows, 96), generator=g, device=dev)
    idx = torch.randperm(rows, generator=g, device=dev)[:n]
    o3, o2, oc = torch.empty((n, 16, 3), device=dev), torch.empty((n, 16, 2), device=dev), torch.empty((n, 9), device=dev)
    for _ in range(reps):
        _cabi.check(lib.dhfk_bank_gather(P(bank), 96, 9, P(idx), n, rows, P(o3), P(o2), P(oc), st), "bank")
elif which == "video":
    F = 9
    n = n // F * F
    x = torch.randn((n, 16, 3), generator=g, device=dev) * 0.4
    b = n // F
    k, dk = torch.empty((b, F, 15), device=dev), torch.empty((b, F - 1, 15), device=dev)
    dp = torch.empty((b, F - 1, 48), device=dev)
    gx = torch.empty((n, 16, 3), device=dev)
    for _ in range(reps):
        _cabi.check(lib.dhfk_video_critic_forward(P(x), F, 0, P(k), P(dk), P(dp), None, n, st), "video fwd")
        _cabi.check(lib.dhfk_video_critic_backward(P(x), F, 0, P(k), P(dk), P(dp), None, P(gx), n, st), "video bwd")
        _cabi.check(lib.dhfk_video_critic_jvp(P(x), P(gx), F, 0, P(k), P(dk), P(dp), None, n, st), "video jvp")
elif which == "wide":
    slots = torch.zeros((n, 37), device=dev)
    slots[:, :33] = d["ang"]; slots[:, 34:] = d["grot"]
    gs = torch.empty((n, 37), device=dev)
    gr = torch.empty((n, 3), device=dev)
    for _ in range(reps):
        _cabi.check(lib.dhfk_forward(P(slots), 37, P(slots) + 136, 37, P(d["bone"]), 15, P(d["root"]), 3, None,
                                     P(world), None, None, n, 0, st), "wide fwd")
        _cabi.check(lib.dhfk_backward(P(slots), 37, P(slots) + 136, 37, P(d["bone"]), 15, P(d["root"]), 3, None,
                                      P(gw), None, None, P(gs), 37, P(gs) + 136, 37, P(gr), 3, None, 15, n, 0, st), "wide bwd")
torch.cuda.synchronize()
print("ok", which, n)
