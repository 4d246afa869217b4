#!/usr/bin/env python
"""Stage the UNMODIFIED reference under baseline/_ref/ so that it travels to the GPU box.

The base contract installs the reference with `pip install --target baseline/_ref /root/reference`.  That was tried
here and fails: the reference is a flat directory of scripts with neither setup.py nor pyproject.toml
("Directory '/root/reference' is not installable").  This script does what that install would have done for the files
the hot path and its callers need: a byte-for-byte copy (verified by sha256, written to baseline/_ref/MANIFEST.json)
of the model-side Python files and the JSON configs.  baseline/_ref/ is git-ignored (reference sources never enter the
history) but not gpurun-ignored, so `bench.py --impl reference`, the `cpu_baseline` / `torch_eager_gpu` legs and the
`-m gpu` drop-in tests run the reference's own code on the box.

    python baseline/stage_ref.py            # (re)stage from /root/reference
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["models.py", "modules.py", "commons.py", "attentions.py", "transforms.py", "stft.py", "pqmf.py", "LICENSE"]


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    if not os.path.isdir(SRC):
        print(f"{SRC} is not present (GPU box?): nothing to stage", file=sys.stderr)
        return 0 if os.path.isdir(DST) else 1
    os.makedirs(os.path.join(DST, "configs"), exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        manifest[f] = sha(os.path.join(DST, f))
        assert manifest[f] == sha(os.path.join(SRC, f))
    for f in sorted(os.listdir(os.path.join(SRC, "configs"))):
        if f.endswith(".json"):
            shutil.copyfile(os.path.join(SRC, "configs", f), os.path.join(DST, "configs", f))
            manifest["configs/" + f] = sha(os.path.join(DST, "configs", f))
    json.dump({"source": SRC, "sha256": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    print(f"staged {len(manifest)} files under {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
