#!/usr/bin/env python
"""A/B helper: run bench.py (device-resident numbers only) against several builds of libdhfk.so.
usage: python tools/ab_bench.py [--fast-trig] name=path/to/lib.so ...   (name 'default' = in-tree lib)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
extra = [a for a in sys.argv[1:] if a.startswith("--")]
for spec in [a for a in sys.argv[1:] if not a.startswith("--")]:
    name, _, path = spec.partition("=")
    env = dict(os.environ)
    if path:
        env["DHFK_LIB_PATH"] = os.path.join(ROOT, path)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "100", "--warmup", "5",
                          "--no-e2e", "--no-cpu-baseline", "--no-extras"] + extra, env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print("%-10s %.3e poses/s  step %.4f ms  fwd %.4f ms (%.1f%%)  bwd %.4f ms (%.1f%%)  clk %s %s" % (
            name, d["value"], d["ms_per_step"], d["roofline_fwd"]["ms_per_launch"], 100 * d["roofline_fwd"]["frac"],
            d["roofline"]["ms_per_launch"], 100 * d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]),
            flush=True)
        sus, gen = d.get("sustained", {}), d.get("generator_mode", {})
        print("%-10s   sustained %.3e poses/s (%.4f ms/step)   generator mode %.4f ms/step (%.3f)" % (
            "", sus.get("poses_per_s", float("nan")), sus.get("ms_per_step", float("nan")),
            gen.get("ms_per_step", float("nan")), gen.get("frac_of_copy_peak", float("nan"))), flush=True)
    except Exception as e:
        print(name, "FAILED", e, out.stdout[-500:], out.stderr[-1500:], flush=True)
