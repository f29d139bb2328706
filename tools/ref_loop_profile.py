#!/usr/bin/env python
"""Where does an iteration of the reference's own GAN loops go once the drop-in is installed?  cProfile (host side) of
oracle/ref_loop.run_loop on the GPU, BASELINE configs[2] / configs[3] sizes.  Test / measurement infrastructure only."""
import cProfile
import io
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402

import ref_loop  # noqa: E402

inst = dict(generators=True, critics=True, loader_refresh=True)
for mode, batch in (("single", 1024), ("video", 512)):
    fd = os.dup(1)
    os.dup2(2, 1)
    ref_loop.run_loop(mode, device="cuda", iters=4, batch=batch, dense=256, seed=5, install=inst, prefer_staged=True)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    ref_loop.run_loop(mode, device="cuda", iters=10, batch=batch, dense=256, seed=6, install=inst, prefer_staged=True)
    torch.cuda.synchronize()
    pr.disable()
    os.dup2(fd, 1)
    s = io.StringIO()
    st = pstats.Stats(pr, stream=s)
    st.sort_stats("tottime").print_stats(32)
    print("==== %s: top by internal time (10 iterations)" % mode)
    print("\n".join(l[:200] for l in s.getvalue().splitlines()[:50]))
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print("==== %s: top by cumulative time" % mode)
    print("\n".join(l[:200] for l in s.getvalue().splitlines()[:64]))
