"""mega_core.utils.checkpoint (utils/checkpoint.py, utils/model_serialization.py of the reference): a checkpoint written
the way the reference's trainer writes it -- {"model": state_dict} with the "module." prefix of DistributedDataParallel
-- loads into the B200 module tree through the reference's longest-suffix key alignment."""
import os

import pytest
import torch


def test_detectron_checkpointer_loads_a_ddp_checkpoint(tmp_path):
    from mega_core.b200 import synth
    from mega_core.modeling.detector import detectors
    from mega_core.utils.checkpoint import DetectronCheckpointer, align_and_update_state_dicts
    sd = synth.make_state_dict("dff_r101_tiny", seed=9)
    cfg = detectors.vid_config("dff", "R-101-C4", "cpu")
    model = detectors.build_detection_model(cfg)
    tiny = {k: v for k, v in model.state_dict().items()}
    for k, v in sd.items():                               # the tiny net is a subset of the R-101 module tree
        assert k in tiny and tiny[k].shape == v.shape, k
    path = os.path.join(tmp_path, "model_final.pth")
    torch.save({"model": {"module." + k: v for k, v in sd.items()}, "optimizer": {"x": 1}, "iteration": 7}, path)
    extra = DetectronCheckpointer(cfg, model, save_dir=str(tmp_path)).load(path, use_latest=False, flownet=None)
    assert extra == {"iteration": 7}
    got = model.state_dict()
    for k, v in sd.items():
        assert torch.equal(got[k], v), k
    # the reference's three-pass rule: flownet=False leaves flownet parameters alone
    state = {k: torch.zeros_like(v) for k, v in model.state_dict().items()}
    align_and_update_state_dicts(state, sd, flownet=False)
    assert state["flownet.conv2.weight"].abs().sum() == 0 and torch.equal(state["rpn.head.conv.weight"], sd["rpn.head.conv.weight"])
    # longest suffix wins: "stem.conv1.weight" must not be taken for "...layer1.0.conv1.weight"
    state = {"backbone.body.layer1.0.conv1.weight": torch.zeros(1), "backbone.body.stem.conv1.weight": torch.zeros(1)}
    align_and_update_state_dicts(state, {"conv1.weight": torch.ones(1), "stem.conv1.weight": torch.full((1,), 2.0)}, flownet=None)
    assert state["backbone.body.layer1.0.conv1.weight"].item() == 1 and state["backbone.body.stem.conv1.weight"].item() == 2
    with pytest.raises(NotImplementedError):
        DetectronCheckpointer(cfg, model).load("catalog://ImageNetPretrained/MSRA/R-101")


def test_c2_pkl_conversion_matches_reference_names(tmp_path):
    """utils/c2_model_loading.load_c2_format: the state_dict names produced for a Detectron-style R-101-C4 blob list equal
    the reference's own renaming (tests/golden/c2_names.pt: the unmodified _rename_basic_resnet_weights + RPN prefix, run
    in a separate process), momentum blobs are dropped, and the result loads into the module tree by suffix alignment"""
    import pickle
    import numpy as np
    from mega_core.modeling.detector import detectors
    from mega_core.utils.checkpoint import DetectronCheckpointer
    from mega_core.utils.c2_model_loading import rename_c2_keys
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2_names.pt"))
    assert dict(zip(gold["names"], rename_c2_keys(gold["names"]))) == gold["reference"]
    cfg = detectors.vid_config("base", "R-101-C4", "cpu")
    model = detectors.build_detection_model(cfg)
    want = model.state_dict()
    blobs, g = {}, np.random.default_rng(0)
    for c2, name in gold["reference"].items():
        target = [k for k in want if k.endswith(name)]
        if "_momentum" in c2 or not target:
            blobs[c2] = np.zeros(3, np.float32)
            continue
        blobs[c2] = g.standard_normal(tuple(want[target[0]].shape)).astype(np.float32)
    path = os.path.join(tmp_path, "R-101.pkl")
    with open(path, "wb") as f:
        pickle.dump({"blobs": {k: v for k, v in blobs.items() if not k.startswith(("fc1000", "pred_", "cls_score", "bbox_pred"))}}, f)
    DetectronCheckpointer(cfg, model).load(path, use_latest=False, flownet=None)
    got = model.state_dict()
    assert torch.equal(got["backbone.body.layer3.22.conv3.weight"], torch.from_numpy(blobs["res4_22_branch2c_w"]))
    assert torch.equal(got["backbone.body.stem.bn1.weight"], torch.from_numpy(blobs["res_conv1_bn_s"]))
    assert torch.equal(got["roi_heads.box.feature_extractor.head.layer4.0.downsample.0.weight"], torch.from_numpy(blobs["res5_0_branch1_w"]))
    assert torch.equal(got["rpn.head.conv.weight"], torch.from_numpy(blobs["conv_rpn_w"]))
