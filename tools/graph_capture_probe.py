#!/usr/bin/env python
"""Why does capturing the unmodified caller's path in a CUDA graph fail at 1 024 poses and work at 4 608?
Captures the step of tools/dropin_path_bench.py at several sizes and in several orders, printing the full error and
which stage of the step broke the capture (each stage captured on its own first)."""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import dropin_path_bench as dp  # noqa: E402


def capture(step, dev, mode):
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, capture_error_mode=mode):
        step()
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        graph.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / 200 * 1e3


def main():
    dev = torch.device("cuda", 0)
    if "--measure" in sys.argv:        # the bench extra's own code path, in its own order (1 M, then 1 024, then 4 608)
        import json
        out = dp.measure(dev, big=1 << 20)
        print(json.dumps({k: out[k] for k in ("at_1024", "at_4608")}, indent=1))
        return
    for n in (1024, 1056, 2048, 4608, 992, 1024):
        for mode in ("thread_local", "relaxed", "global"):
            step = dp._setup(n, dev)
            for _ in range(10):
                step()
            torch.cuda.synchronize(dev)
            try:
                us = capture(step, dev, mode)
                print("n=%d mode=%s: %.2f us per replay" % (n, mode, us), flush=True)
            except Exception:
                print("n=%d mode=%s: FAILED" % (n, mode), flush=True)
                traceback.print_exc(limit=6)
                try:
                    torch.cuda.synchronize(dev)
                except Exception as e:
                    print("  sync after failure:", repr(e)[:200])
            del step


if __name__ == "__main__":
    main()
