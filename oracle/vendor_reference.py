"""Recipe for `oracle/_ref/`: the UNMODIFIED reference modules of the hot path, made available where /root/reference does
not exist (the GPU box).  Test / baseline infrastructure only (see oracle/__init__.py).

The reference is pure Python with no build or install metadata (no setup.py / pyproject), so "building" it is a byte copy
of the files the path needs -- `src/__init__.py`, `src/models/*`, `src/loss/*`, `src/utils/*`, `src/datasets/*` (the dataset
classes this repo's `src.datasets` re-exports), and the entry scripts `scripts/estimate.py` and `scripts/train_and_evaluate.py`
whose flows tests/test_gpu_dropin.py executes against this repo's `src` -- from where they lie under
/root/reference into `oracle/_ref/`, plus a MANIFEST.json with each file's sha256.  `oracle/_ref/` is git-ignored (no
reference source enters the history) but travels with `gpurun`.  `__graft_entry__.build()` runs this when /root/reference
is present; `bench.py --impl reference` and the `gpu_baseline` leg import the copy (`kind: "reference"`) and fall back to
the oracle port (`kind: "port"`) when it is absent.

    python oracle/vendor_reference.py [/root/reference]
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
PARTS = ("__init__.py", "models", "loss", "utils", "datasets")
SCRIPTS = ("scripts/estimate.py", "scripts/train_and_evaluate.py")


def vendor(reference_root: str = "/root/reference") -> bool:
    src = os.path.join(reference_root, "src")
    if not os.path.isdir(src):
        return False
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(os.path.join(DEST, "src"))
    manifest = {}
    for part in PARTS:
        p = os.path.join(src, part)
        files = [p] if os.path.isfile(p) else [os.path.join(p, f) for f in sorted(os.listdir(p)) if f.endswith(".py")]
        for f in files:
            rel = os.path.relpath(f, reference_root)
            out = os.path.join(DEST, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(f, out)
            with open(f, "rb") as fh:
                manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    for rel in SCRIPTS:
        f = os.path.join(reference_root, rel)
        if os.path.isfile(f):
            out = os.path.join(DEST, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(f, out)
            with open(f, "rb") as fh:
                manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": reference_root, "files": manifest}, fh, indent=1)
    return True


def verify() -> bool:
    """True when every vendored file still has the recorded sha256 (i.e. is the unmodified reference file)."""
    try:
        with open(os.path.join(DEST, "MANIFEST.json")) as fh:
            files = json.load(fh)["files"]
        for rel, digest in files.items():
            with open(os.path.join(DEST, rel), "rb") as fh:
                if hashlib.sha256(fh.read()).hexdigest() != digest:
                    return False
        return bool(files)
    except (OSError, ValueError, KeyError):
        return False


if __name__ == "__main__":
    ok = vendor(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("vendored" if ok else "reference not present; nothing done")
