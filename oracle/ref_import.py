"""Import the UNMODIFIED reference (`/root/reference/mega_core`) in this container, behind shims.

Used only by oracle/make_golden.py and the here-only validation tests to (a) pin the oracle
restatement against the reference itself and (b) generate the fixtures in tests/golden/.
`/root/reference` does not exist on the GPU box: nothing on the GPU-side path imports this.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("MEGA_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE, "mega_core"))


def setup():
    """returns the reference's `mega_core` package, imported from where it lies"""
    if not available():
        raise ImportError("reference checkout not present at %s" % REFERENCE)
    import numpy as np
    import torch
    if not hasattr(np, "float"):
        np.float = float  # anchor_generator.py:229-238
    if not hasattr(np, "bool"):
        np.bool = bool
    if not hasattr(torch, "_six"):
        torch._six = types.SimpleNamespace(PY3=True, string_classes=(str,))
    shims = os.path.join(HERE, "shims")
    if shims not in sys.path:
        sys.path.insert(0, shims)
    os.environ["PATH"] = os.path.join(shims, "bin") + os.pathsep + os.environ.get("PATH", "")
    for m in ("pycocotools", "pycocotools.coco", "pycocotools.mask", "pycocotools.cocoeval",
              "cityscapesscripts", "cityscapesscripts.helpers", "cityscapesscripts.helpers.csHelpers"):
        sys.modules.setdefault(m, types.ModuleType(m))
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import build_ref
    build_ref.build()
    # the package must exist before `mega_core._C` can be registered under it
    import importlib
    pkg = importlib.import_module("mega_core")
    assert pkg.__file__.startswith(REFERENCE), "a different mega_core is already imported: %s" % pkg.__file__
    sys.modules["mega_core._C"] = build_ref.load_module()
    pkg._C = sys.modules["mega_core._C"]
    return pkg


def build_cfg(config_file, opts=()):
    """cfg exactly as tools/test_net.py:75-79 builds it (BASE_RCNN_1gpu.yaml, method yaml, opts)"""
    setup()
    from mega_core.config import cfg
    c = cfg.clone()
    c.merge_from_file(os.path.join(REFERENCE, "configs", "BASE_RCNN_1gpu.yaml"))
    c.merge_from_file(os.path.join(REFERENCE, config_file))
    c.merge_from_list(["MODEL.DEVICE", "cpu"] + list(opts))
    c.freeze()
    return c
