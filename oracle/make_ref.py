#!/usr/bin/env python3
"""Recipe for `oracle/_ref/`: a verbatim, git-ignored copy of the reference's Python package so that the UNMODIFIED
reference can be imported on the GPU box (which has no /root/reference) - as the timed `--impl reference` CPU arm and
the `--impl eager` GPU-eager arm of bench.py, and as a second checker beside the oracle port.

TEST / BENCH INFRASTRUCTURE ONLY.  Nothing is copied into the repository history (`oracle/_ref/` is listed in
.gitignore, not in .gpurunignore, so the snapshot that travels to the GPU box carries it).  Run by
`__graft_entry__.build()` whenever /root/reference is present:

    python oracle/make_ref.py
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("FLAMED_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def make_ref(force=False):
    if not os.path.isdir(os.path.join(SRC, "flamed")):
        return None  # not in the build container: keep whatever copy is already there
    stamp = os.path.join(DST, ".copied_from")
    if not force and os.path.exists(stamp) and os.path.isdir(os.path.join(DST, "flamed")):
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", "*.bin", "*.pt", "*.ckpt")
    shutil.copytree(os.path.join(SRC, "flamed"), os.path.join(DST, "flamed"), ignore=ignore)
    shutil.copytree(os.path.join(SRC, "configs"), os.path.join(DST, "configs"), ignore=ignore)
    with open(stamp, "w") as f:
        f.write(SRC + "\n")
    return DST


if __name__ == "__main__":
    print(make_ref(force="--force" in sys.argv))
