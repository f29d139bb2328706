#!/usr/bin/env python
"""How far each trig variant of the kernels sits from (a) exact arithmetic (float64 C oracle) and (b) the reference's
own fp32 results (torch port, bit-identical to the reference), next to the reference's own distance from exact
arithmetic.  Metric: max |x - ref| / max(|ref|, 1) per quantity (the north-star tolerance is 1e-5)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import dhfk, c_oracle, torch_port
from dhfk import synthetic, tables

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
inp = synthetic.gan_like(n, seed=1234); up = synthetic.upstream_grads(n, seed=4321)
blk = tables.camera_block("S1", 0)
rel = lambda x, r: float((np.abs(np.asarray(x, np.float64) - r) / np.maximum(np.abs(r), 1.0)).max())
o = c_oracle.forward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk)
b = c_oracle.backward(inp["ang"], inp["grot"], inp["bone"], inp["root"], blk, g_world=up["g_world"], g_uv=up["g_uv"], want_bone=False)
exact = dict(world=o["world16"], uv=o["uv"], g_ang=b["g_ang"], g_grot=b["g_grot"], g_root=b["g_root"])
T = lambda a, g=False: torch.tensor(a, device="cuda:0", requires_grad=g)
def run(kw):
    a, g, r = T(inp["ang"], True), T(inp["grot"], True), T(inp["root"], True)
    w, _, uv = dhfk.fk_project(a, g, T(inp["bone"]), r, blk, return_cam=False, **kw)
    ((w * T(up["g_world"])).sum() + (uv * T(up["g_uv"])).sum()).backward()
    return dict(world=w.detach().cpu().numpy(), uv=uv.detach().cpu().numpy(), g_ang=a.grad.cpu().numpy(),
                g_grot=g.grad.cpu().numpy(), g_root=r.grad.cpu().numpy())
ref = {k: [] for k in exact}
torch.set_num_threads(os.cpu_count())
for lo in range(0, n, 1024):
    sl = slice(lo, lo + 1024)
    a = torch.tensor(inp["ang"][sl], requires_grad=True); g = torch.tensor(inp["grot"][sl], requires_grad=True)
    r = torch.tensor(inp["root"][sl], requires_grad=True)
    _, w16, _, uv = torch_port.pipeline(a, g, torch.tensor(inp["bone"][sl]), r, blk)
    ((w16 * torch.tensor(up["g_world"][sl])).sum() + (uv * torch.tensor(up["g_uv"][sl])).sum()).backward()
    for k, v in (("world", w16), ("uv", uv), ("g_ang", a.grad), ("g_grot", g.grad), ("g_root", r.grad)):
        ref[k].append(v.detach().numpy())
ref = {k: np.concatenate(v).astype(np.float64) for k, v in ref.items()}
print("n = %d poses; max |x-ref|/max(|ref|,1)" % n)
print("%-34s" % "" + "".join("%10s" % k for k in exact))
print("%-34s" % "reference fp32  vs exact" + "".join("%10.2e" % rel(ref[k], exact[k]) for k in exact))
for name, kw in (("kernels ACCURATE_TRIG flag", dict(accurate_grad=True)), ("kernels default", {}),
                 ("kernels FAST_TRIG flag", dict(fast_trig=True))):
    got = run(kw)
    print("%-34s" % (name + " vs exact") + "".join("%10.2e" % rel(got[k], exact[k]) for k in exact))
    print("%-34s" % (name + " vs reference fp32") + "".join("%10.2e" % rel(got[k], ref[k]) for k in exact))
