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
