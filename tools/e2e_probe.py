#!/usr/bin/env python
"""Probe: raw pinned-memory copy bandwidth of the box and the e2e path for several chunk sizes / stream counts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, dhfk
from dhfk import synthetic, tables

dev = torch.device("cuda", 0)
n = 1 << 20
buf_h = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True); buf_d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
buf_h2 = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True); buf_d2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def bw(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return reps * 0.256 * 1.048576 / (time.perf_counter() - t)
print("H2D alone  %.1f GB/s" % bw(lambda: buf_d.copy_(buf_h, non_blocking=True)))
print("D2H alone  %.1f GB/s" % bw(lambda: buf_h.copy_(buf_d, non_blocking=True)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): buf_d.copy_(buf_h, non_blocking=True)
    with torch.cuda.stream(s2): buf_h2.copy_(buf_d2, non_blocking=True)
print("H2D+D2H concurrent: %.1f GB/s each direction" % bw(both))
hin = synthetic.gan_like(1 << 16, seed=5); up = synthetic.upstream_grads(1 << 16, seed=6)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(np.tile(a, (16,) + (1,) * (a.ndim - 1)))).pin_memory()
h = [pin(hin[k]) for k in ("ang", "grot", "bone", "root")]; gw, gu = pin(up["g_world"]), pin(up["g_uv"])
blk = tables.camera_block("S1", 0)
for chunk in (1 << 15, 1 << 16, 1 << 17, 1 << 18):
    for ns in (2, 3, 4, 6):
        out = {}
        for _ in range(2):
            out = dhfk.fk_project_host(*h, blk, gw, gu, chunk_rows=chunk, num_streams=ns, workspace=out.get("_workspace"), out=out)
        t = time.perf_counter()
        for _ in range(5):
            out = dhfk.fk_project_host(*h, blk, gw, gu, chunk_rows=chunk, num_streams=ns, workspace=out["_workspace"], out=out)
        dt = (time.perf_counter() - t) / 5
        print("chunk %7d streams %d: %.2f ms  %.3e poses/s" % (chunk, ns, dt * 1e3, n / dt), flush=True)
