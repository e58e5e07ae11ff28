#!/usr/bin/env python
"""Stage the reference's hot-path modules into oracle/_ref/ (git-ignored, NOT gpurun-ignored).

TEST / BENCH INFRASTRUCTURE.  The reference is two pure-Python files on this path — nothing to
compile — so "building" oracle/_ref means copying `/root/reference/{lsh,recommender}.py` there byte
for byte (sha256 recorded in MANIFEST.json) so that `bench.py --impl reference` and the
`cpu_baseline` leg can time the UNMODIFIED reference on the GPU box, where /root/reference does
not exist.  The files are build outputs: they never enter the git history (.gitignore) and no
product module imports them.  Run by `__graft_entry__.build()` whenever /root/reference is present.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
FILES = ("lsh.py", "recommender.py", "requirements.txt")


def stage(ref_root: str = REF) -> bool:
    if not os.path.isdir(ref_root):
        return False
    out = os.path.join(HERE, "_ref")
    os.makedirs(out, exist_ok=True)
    manifest = {}
    for name in FILES:
        src = os.path.join(ref_root, name)
        shutil.copyfile(src, os.path.join(out, name))
        manifest[name] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(out, "MANIFEST.json"), "w") as fh:
        json.dump({"source": ref_root, "sha256": manifest,
                   "note": "byte-for-byte copies of the reference's hot-path modules; build output, git-ignored"}, fh, indent=1)
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged oracle/_ref" if ok else "no /root/reference here: nothing staged")
    sys.exit(0)
