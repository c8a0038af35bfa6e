"""CPU tests of the C-ABI boundary: the library loads without a GPU and exports every symbol that
include/mega_b200.h declares; argument validation that does not need a device; no CPU fallback."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mega_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mega_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mega_core import _lib
    syms = _declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(_lib.lib, s), "libmega_b200.so does not export %s" % s
    assert sorted(_lib.EXPORTS) == syms
    assert _lib.lib.mega_abi_version() == 5


def test_no_cpu_fallback():
    from mega_core import _lib
    from mega_core.b200 import ops
    with pytest.raises(_lib.MegaError):
        ops.nms_device(torch.zeros(4, 4), torch.zeros(4), 0.5)
    with pytest.raises(_lib.MegaError):
        ops.roi_align_nchw(torch.zeros(1, 4, 8, 8), torch.zeros(2, 5), 1 / 16, 7, 7, 0)


def test_argument_validation_without_device():
    from mega_core import _lib
    lib = _lib.lib
    assert lib.mega_nms_workspace_bytes(9000) == -1
    assert lib.mega_nms_workspace_bytes(6000) > 6000 * 94 * 8
    assert lib.mega_rpn_select_workspace_bytes(2, 38, 63, 12, 6000) > 0
    assert lib.mega_rpn_select_workspace_bytes(1, 38, 63, 12, 9000) == -1
    assert lib.mega_box_postprocess_workspace_bytes(300, 31) > 0
    assert lib.mega_box_postprocess_workspace_bytes(600, 31) == -1
    d = _lib.ConvGemmDesc()
    d.tile_h, d.tile_w, d.block_n, d.batch = 3, 40, 64, 1
    assert lib.mega_conv_gemm_tf32(ctypes.byref(d), None) != 0
    assert b"tile_h*tile_w" in lib.mega_last_error()


def test_tile_and_block_heuristics():
    from mega_core.b200 import ops
    for h, w in ((38, 63), (150, 250), (75, 125), (1, 375), (12, 20)):
        th, tw = ops.pick_tile(h, w)
        assert th * tw == 128
    assert ops.pick_config(60, 38, 1, 32)[0] in (32, 64) and ops.pick_config(1024, 38, 1, 288) == (256, 1)
    assert ops.pick_config(256, 38, 1, 32) == (96, 0)


def test_engine_tables_host_logic():
    """ring-buffer index tables of the MEGA engine (pure host logic, CPU tensors)"""
    from mega_core.b200 import engine
    eng = engine.MegaEngine.__new__(engine.MegaEngine)
    c = engine.EngineConfig()
    assert c.advanced_num == 15
    ca = engine.cell_anchors(16, c.anchor_sizes, c.aspect_ratios)
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "reference_ops.pt"))
    assert torch.equal(ca, gold["cell_anchors"])   # the reference's generate_anchors output
