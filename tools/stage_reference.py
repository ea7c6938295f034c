#!/usr/bin/env python
"""Stage the UNMODIFIED reference into baseline/_ref/ (git-ignored; it travels to the GPU box with gpurun).

    python tools/stage_reference.py            # needs /root/reference (read-only), writes baseline/_ref/

The reference is pure Python without setup.py / pyproject, so `pip install /root/reference` cannot work; its
"install" is a verbatim copy of the source files the hot path and its callers need (model/, engine/, utils/,
tools/, config/, train.py, test.py).  Nothing is edited, and nothing from baseline/_ref/ is tracked by git or
imported by the product (`cris/`): it is driven only by `bench.py --impl reference`, bench.py's
`gpu_incumbent` / `cpu_baseline` legs and `tests/test_dropin_gpu.py`, as the thing measured against /
the caller side of the drop-in boundary.  A manifest with the sha256 of every staged file is written next to it,
so a test can prove the staged copy is byte-identical to the source it came from.
"""
import hashlib
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(REPO, "baseline", "_ref")
WHAT = ["model", "engine", "utils", "tools", "config", "train.py", "test.py", "LICENSE"]
SKIP_SUFFIX = (".pyc", ".gif", ".png")


def stage(src: str = SRC, dst: str = DST) -> dict:
    if not os.path.isdir(src):
        raise FileNotFoundError(f"{src} is not available here (it only exists in the build container)")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    manifest = {}
    for item in WHAT:
        s = os.path.join(src, item)
        if not os.path.exists(s):
            continue
        files = [s] if os.path.isfile(s) else [os.path.join(r, f) for r, _, fs in os.walk(s) for f in fs]
        for f in sorted(files):
            if f.endswith(SKIP_SUFFIX) or "__pycache__" in f:
                continue
            rel = os.path.relpath(f, src)
            out = os.path.join(dst, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(f, out)
            manifest[rel] = hashlib.sha256(open(out, "rb").read()).hexdigest()
    json.dump({"source": src, "files": manifest}, open(os.path.join(dst, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    m = stage()
    print(f"[stage_reference] {len(m)} files -> {DST}", file=sys.stderr)
