#!/usr/bin/env python
"""Stage the UNMODIFIED reference files of the hot path under baseline/_ref/ (git-ignored, NOT gpurun-ignored, so the copy
travels to the GPU box where /root/reference does not exist).  The reference is a directory of scripts without a
setup.py / pyproject, so `pip install --target baseline/_ref /root/reference` is not applicable; the three pure-Python files
that make up the path are copied byte for byte instead:
    src/deepCam/architecture/{__init__,deeplab_xception}.py   (DeepLabv3_plus and its building blocks)
    src/deepCam/utils/losses.py                               (fp_loss)
Only bench.py's measurement arms (`--impl reference`: the reference's own training step on the host cores;
`--impl torch_gpu`: the same classes on the GPU through torch/cuDNN as the library anchor) load them, by file path under
private module names; nothing in mlperf-deepcam_b200/ can import them.  Called by __graft_entry__.build() whenever
/root/reference is present."""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DEEPCAM_REFERENCE", "/root/reference/src/deepCam")
DST = os.path.join(HERE, "_ref", "deepCam")
FILES = ["architecture/__init__.py", "architecture/deeplab_xception.py", "utils/losses.py"]


def stage(verbose=False):
    """Returns True when baseline/_ref holds the files (copied now or already identical), False when there is no reference."""
    if not os.path.isdir(SRC):
        return all(os.path.exists(os.path.join(DST, f)) for f in FILES)
    for f in FILES:
        s, d = os.path.join(SRC, f), os.path.join(DST, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
            if verbose:
                print("staged", d)
    return True


if __name__ == "__main__":
    ok = stage(verbose=True)
    print("baseline/_ref:", "ready" if ok else "no reference checkout at %s" % SRC)
    sys.exit(0 if ok else 1)
