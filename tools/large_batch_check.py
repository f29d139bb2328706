import sys, numpy as np, torch
sys.path.insert(0,'.'); sys.path.insert(0,'oracle')
import dhfk, c_oracle
from dhfk import synthetic, tables
n = 23_000_017
dev = torch.device('cuda:0')
d = synthetic.gan_like_torch(n, dev, seed=3)
blk = tables.camera_block('S5', 2)
a, g, r = d['ang'].requires_grad_(True), d['grot'].requires_grad_(True), d['root'].requires_grad_(True)
w, _, uv = dhfk.fk_project(a, g, d['bone'], r, blk, return_cam=False)
gw = torch.randn(n, 16, 3, device=dev); gu = torch.randn(n, 16, 2, device=dev)
((w * gw).sum() + (uv * gu).sum()).backward()
worst = 0
for lo in (0, n - 4096, (1 << 32) // 192 - 2048, (1 << 31) // 192 - 2048, (1 << 31) // 132 - 2048, (1 << 32) // 128 - 2048 if (1 << 32) // 128 < n else 1 << 20):
    sl = slice(lo, lo + 4096)
    c = lambda t: t[sl].detach().cpu().numpy()
    o = c_oracle.forward(c(a), c(g), c(d['bone']), c(r), blk)
    b = c_oracle.backward(c(a), c(g), c(d['bone']), c(r), blk, g_world=c(gw), g_uv=c(gu), want_bone=False)
    rel = lambda x, ref: float((np.abs(x - ref) / np.maximum(np.abs(ref), 1)).max())
    errs = (rel(c(w), o['world16']), rel(c(uv), o['uv']), rel(c(a.grad), b['g_ang']), rel(c(g.grad), b['g_grot']), rel(c(r.grad), b['g_root']))
    worst = max(worst, *errs)
    print(lo, ' '.join('%.2e' % e for e in errs))
assert worst < 1e-5
print('ok', n)
