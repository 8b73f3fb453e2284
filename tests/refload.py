"""Import shim for the read-only reference checkout (only available in the build container).

Nothing under -m gpu, smoke() or bench.py may use this; it backs the `reference`-marked tests and
tests/golden/make_golden.py.
"""
import importlib.util
import os
import sys
import types

REFERENCE = "/root/reference/src/deepCam"


def available():
    return os.path.isdir(REFERENCE)


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE, relpath))
    mod = importlib.util.module_from_spec(spec)
    old = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.dont_write_bytecode = old
    return mod


def deeplab():
    return _load("_ref_deeplab_xception", "architecture/deeplab_xception.py")


def losses():
    return _load("_ref_losses", "utils/losses.py")


def utils():
    # utils/utils.py imports matplotlib.pyplot (UT:29) which is absent here; it is unused by compute_score
    if "matplotlib" not in sys.modules:
        m = types.ModuleType("matplotlib")
        m.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.pyplot"] = m.pyplot
    return _load("_ref_utils", "utils/utils.py")
