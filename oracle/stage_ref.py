#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- stage the unmodified reference so that it can run on the GPU box.

    python oracle/stage_ref.py            # /root/reference/DH-AUG_master  ->  oracle/_ref/dh_aug_ref.zip

The reference is plain Python + torch: there is nothing to compile.  What the "build it into oracle/_ref" recipe
amounts to here is one zip archive of its importable packages (common, models_Fk_GAN, utils, function_aug,
models_baseline, progress) plus the bone-length templates the loader refresh reads, byte for byte as they lie under
/root/reference, imported straight from the archive (zipimport) by oracle/ref_harness.py.  `oracle/_ref/` is
git-ignored (no reference source enters the history) but not gpurun-ignored, so the archive travels to the GPU box,
where /root/reference does not exist.  Used by: tests (`-m gpu`: the reference's own GAN loops with the drop-in
installed), bench.py's `--impl reference` arm and `cpu_baseline` leg (kind "reference").  Never by the product.
__graft_entry__.build() calls this when /root/reference is present.
"""
from __future__ import annotations

import hashlib
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DHFK_REFERENCE_ROOT", "/root/reference/DH-AUG_master")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "dh_aug_ref.zip")
PACKAGES = ("common", "models_Fk_GAN", "utils", "function_aug", "models_baseline", "progress")
EXTRA = ("data_extra/bone_length_npy",)


def _files():
    for pkg in PACKAGES:
        for root, _dirs, names in os.walk(os.path.join(SRC, pkg)):
            for n in sorted(names):
                if n.endswith(".py"):
                    yield os.path.join(root, n)
    for sub in EXTRA:
        for root, _dirs, names in os.walk(os.path.join(SRC, sub)):
            for n in sorted(names):
                yield os.path.join(root, n)


def stage(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "models_Fk_GAN")):
        raise SystemExit("reference tree not found at %s" % SRC)
    os.makedirs(OUT_DIR, exist_ok=True)
    files = sorted(_files())
    h = hashlib.sha256()
    tmp = OUT + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        dirs = set()
        for f in files:
            d = os.path.dirname(os.path.relpath(f, SRC))
            while d:
                dirs.add(d)
                d = os.path.dirname(d)
        for d in sorted(dirs):    # explicit directory entries: zipimport needs them for packages without __init__.py
            z.writestr(zipfile.ZipInfo(d + "/", date_time=(2020, 1, 1, 0, 0, 0)), b"")
        for f in files:
            rel = os.path.relpath(f, SRC)
            data = open(f, "rb").read()
            h.update(rel.encode()); h.update(data)
            # fixed timestamps: the archive is reproducible byte for byte
            z.writestr(zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0)), data, zipfile.ZIP_DEFLATED)
    os.replace(tmp, OUT)
    with open(os.path.join(OUT_DIR, "MANIFEST.txt"), "w") as m:
        m.write("source %s\nfiles %d\nsha256 %s\n" % (SRC, len(files), h.hexdigest()))
    if verbose:
        print("[stage_ref] %d files -> %s (%d bytes, sha256 %s)" % (len(files), OUT, os.path.getsize(OUT), h.hexdigest()[:16]))
    return OUT


if __name__ == "__main__":
    stage()
    sys.exit(0)
