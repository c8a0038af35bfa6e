"""Caffe2 / Detectron `.pkl` backbone weights -> this package's state_dict names (utils/c2_model_loading.py:11-206 of the
reference; `DetectronCheckpointer._load_file` routes `*.pkl` here, e.g. the ImageNet-pretrained MSRA R-50 / R-101 that
the VID training configs start from). C4 / C5 ResNet bodies only -- the FPN, mask, keypoint and group-norm branches of
the reference's table belong to model families that are out of scope.

The conversion is DATA: an ordered list of substring rewrites applied to every blob name (the order matters: later
rules consume what earlier ones produce), then the RPN prefix and the optional deformable-conv re-nesting. The rule
table below is the subset of the reference's that a C4 / C5 body can trigger; tests/test_checkpoint_cpu.py checks the
resulting names against the reference's own function on a Detectron-style R-101 blob list."""
import logging
import pickle
import re
from collections import OrderedDict

import torch

_RULES = (
    ("_", "."), (".w", ".weight"), (".bn", "_bn"), (".b", ".bias"), ("_bn.s", "_bn.scale"),
    (".biasranch", ".branch"),                       # ".branch..." was hit by the ".b" rule: undo
    ("bbox.pred", "bbox_pred"), ("cls.score", "cls_score"), ("res.conv1_", "conv1_"),
    (".biasbox", ".bbox"), ("conv.rpn", "rpn.conv"), ("rpn.bbox.pred", "rpn.bbox_pred"),
    ("rpn.cls.logits", "rpn.cls_logits"),
    ("_bn.scale", "_bn.weight"),                     # AffineChannel -> (frozen) batch norm naming
    ("conv1_bn.", "bn1."),
    ("res2.", "layer1."), ("res3.", "layer2."), ("res4.", "layer3."), ("res5.", "layer4."),
    (".branch2a.", ".conv1."), (".branch2a_bn.", ".bn1."), (".branch2b.", ".conv2."), (".branch2b_bn.", ".bn2."),
    (".branch2c.", ".conv3."), (".branch2c_bn.", ".bn3."),
    (".branch1.", ".downsample.0."), (".branch1_bn.", ".downsample.1."),
    ("rpn.", "rpn.head."),
)


def rename_c2_keys(names):
    """blob names -> state_dict names, in the given order"""
    out = []
    for k in names:
        k = {"pred_b": "fc1000_b", "pred_w": "fc1000_w"}.get(k, k)          # X-101 classifier blobs
        for old, new in _RULES:
            k = k.replace(old, new)
        out.append(k)
    return out


def _nest_dcn_convs(state_dict, stage_with_dcn):
    """stages with deformable convolutions keep the 3x3 weights one level deeper (conv2.conv.*)"""
    for ix, with_dcn in enumerate(stage_with_dcn, 1):
        if not with_dcn:
            continue
        for key in sorted(state_dict.keys()):
            if re.match(".*layer%d.*conv2.*" % ix, key):
                for param in ("weight", "bias"):
                    if param in key:
                        state_dict[key.replace("conv2.%s" % param, "conv2.conv.%s" % param)] = state_dict.pop(key)
    return state_dict


def load_c2_format(cfg, f):
    body = cfg.MODEL.BACKBONE.CONV_BODY
    if not re.fullmatch(r"R-(50|101|152)-C[45]", body):
        raise NotImplementedError("load_c2_format (B200 build): C4 / C5 ResNet bodies only, got %s" % body)
    with open(f, "rb") as fh:
        data = pickle.load(fh, encoding="latin1")
    blobs = data["blobs"] if "blobs" in data else data
    names = sorted(k for k in blobs.keys())
    logger = logging.getLogger(__name__)
    state = OrderedDict()
    for old, new in zip(names, rename_c2_keys(names)):
        if "_momentum" in old:
            continue
        logger.info("C2 name: %s mapped name: %s", old, new)
        state[new] = torch.from_numpy(blobs[old])
    return dict(model=_nest_dcn_convs(state, cfg.MODEL.RESNETS.STAGE_WITH_DCN))
