#!/usr/bin/env python
"""Per-kernel executed-instruction mix and shared-memory conflict hot spots from an ncu report's source page (SASS view).
usage: python tools/ncu_opcode_mix.py gpurun_out/<tag>_prof.ncu-rep [top]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kernel, hdr = None, None
mix, conf, total, warps = {}, {}, {}, {}
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        kernel = row[1]
        mix[kernel], conf[kernel], total[kernel] = collections.Counter(), [], 0
        hdr = None
        continue
    if row[0] == "Address":
        hdr = {h: i for i, h in enumerate(row)}
        continue
    if hdr is None or kernel is None:
        continue
    src = row[hdr["Source"]].strip()
    n = int(row[hdr["Instructions Executed"]] or 0)
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.rstrip(";")
    mix[kernel][op] += n
    total[kernel] += n
    warps.setdefault(kernel, n)
    if "L1 Wavefronts Shared Excessive" in hdr:       # absent for kernels that never touch shared memory
        ex = int(row[hdr["L1 Wavefronts Shared Excessive"]] or 0)
        if ex:
            conf[kernel].append((ex, int(row[hdr["L1 Wavefronts Shared"]] or 0), src))
for k in mix:
    w = max(warps.get(k, 1), 1)
    print("=== %s\n    executed warp-instructions %d  (%.0f per warp, %d warps)" % (k, total[k], total[k] / w, w))
    for op, n in mix[k].most_common(top):
        print("    %-28s %12d  %6.1f /warp  %5.1f %%" % (op, n, n / w, 100.0 * n / total[k]))
    if conf[k]:
        print("    -- shared-memory wavefronts in excess of ideal, by instruction")
        for ex, wf, src in sorted(conf[k], reverse=True)[:12]:
            print("    %10d of %10d  %s" % (ex, wf, src))
