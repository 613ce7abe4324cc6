#!/usr/bin/env python
"""Stage the UNMODIFIED reference model graphs next to the repo so that they travel to the GPU box.

    python tools/fetch_ref.py [--src /root/reference] [--force]

Copies ``<src>/models/**/*.py`` (vit_quant.py, swin_quant.py, layers_quant.py, model_utils.py, utils.py, __init__.py
and the reference's own quantization_utils/) plus ``TVM_benchmark/convert_model.py`` byte for byte into the git-ignored
``baseline/_ref/`` tree and writes a manifest with the sha256 of every file.  Nothing under ``baseline/_ref`` is product
source or enters history (.gitignore); it is NOT gpurun-ignored, so it ships with the snapshot like the built .so.

Used by
  * tests/test_zz_reference_graphs_gpu.py -- the reference's vit_quant.py / swin_quant.py running unchanged on the sm_100a
    operator mirror (the reference's quantization_utils is replaced through ``ivit_b200.dropin``),
  * bench.py's CPU arm -- the literal reference ``model(x)`` (its own quantization_utils, fp32 carrier) on the host cores.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
WANT = ["models", os.path.join("TVM_benchmark", "convert_model.py")]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def fetch(src: str = "/root/reference", force: bool = False) -> dict:
    """Copy the files; returns the manifest {relative path: sha256}.  No-op (returns the stored manifest) when the
    reference checkout is absent and a staged copy already exists."""
    man_path = os.path.join(DST, "MANIFEST.json")
    if not os.path.isdir(os.path.join(src, "models")):
        if os.path.exists(man_path):
            return json.load(open(man_path))
        raise FileNotFoundError("reference checkout not found at %s and nothing staged under %s" % (src, DST))
    files = []
    for w in WANT:
        p = os.path.join(src, w)
        if os.path.isdir(p):
            for d, _, names in os.walk(p):
                files += [os.path.relpath(os.path.join(d, n), src) for n in names if n.endswith(".py")]
        elif os.path.exists(p):
            files.append(w)
    manifest = {}
    for rel in sorted(files):
        s, t = os.path.join(src, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(t), exist_ok=True)
        if force or not os.path.exists(t) or _sha(t) != _sha(s):
            shutil.copyfile(s, t)
        manifest[rel] = _sha(t)
    with open(man_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.environ.get("IVIT_REFERENCE", "/root/reference"))
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    m = fetch(a.src, a.force)
    print("staged %d reference files under %s" % (len(m), DST))
    sys.exit(0)
