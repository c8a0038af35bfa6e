"""Checkpoint loading with the reference's names (utils/checkpoint.py:13-148, utils/model_serialization.py:10-89):
`DetectronCheckpointer(cfg, model, save_dir=...).load(cfg.MODEL.WEIGHT)` as called by tools/test_net.py:98-100.

Inference side only: native `.pth` checkpoints (the released MEGA / RDN / FGFA / DFF / base models: a dict with a
"model" state_dict, or a bare state_dict), with the reference's key alignment -- a "module." prefix of a
(Distributed)DataParallel save is stripped, then every model key takes the loaded key that is its LONGEST SUFFIX
(so a checkpoint saved under extra or missing name prefixes still lands on the right parameters). Caffe2 `.pkl`
ImageNet backbones go through `c2_model_loading.load_c2_format` (C4 / C5 ResNet bodies); catalog / URL resolution and
optimizer / scheduler state are training-time concerns (the former raises a clear error, the latter is dropped). The module classes of mega_core.modeling build their B200 engine lazily from the
module's state_dict, so weights loaded this way are what the kernels use."""
import logging
import os
from collections import OrderedDict

import torch


def strip_prefix_if_present(state_dict, prefix):
    if not state_dict or not all(k.startswith(prefix) for k in state_dict):
        return state_dict
    return OrderedDict((k.replace(prefix, ""), v) for k, v in state_dict.items())


def align_and_update_state_dicts(model_state_dict, loaded_state_dict, flownet=False):
    """model key <- the loaded key that is its longest suffix. `flownet` selects the reference's three passes
    (model_serialization.py:27-38): False skips flownet / embednet parameters, True takes only flownet ones, None all."""
    loaded = sorted(loaded_state_dict.keys())
    logger = logging.getLogger(__name__)
    for key in sorted(model_state_dict.keys()):
        if flownet is False and ("flownet" in key or "embednet" in key):
            continue
        if flownet is True and "flownet" not in key:
            continue
        best = ""
        for cand in loaded:                               # ascending order + strict ">" = the reference's argmax tie rule
            if key.endswith(cand) and len(cand) > len(best):
                best = cand
        if best:
            model_state_dict[key] = loaded_state_dict[best]
            logger.info("%s loaded from %s of shape %s", key, best, tuple(loaded_state_dict[best].shape))


def load_state_dict(model, loaded_state_dict, flownet=False):
    state = model.state_dict()
    align_and_update_state_dicts(state, strip_prefix_if_present(loaded_state_dict, "module."), flownet=flownet)
    model.load_state_dict(state)


class Checkpointer(object):
    def __init__(self, model, optimizer=None, scheduler=None, save_dir="", save_to_disk=None, logger=None):
        self.model, self.optimizer, self.scheduler = model, optimizer, scheduler
        self.save_dir, self.save_to_disk = save_dir, save_to_disk
        self.logger = logger or logging.getLogger(__name__)

    def has_checkpoint(self):
        return bool(self.save_dir) and os.path.exists(os.path.join(self.save_dir, "last_checkpoint"))

    def get_checkpoint_file(self):
        try:
            with open(os.path.join(self.save_dir, "last_checkpoint"), "r") as f:
                return f.read().strip()
        except IOError:
            return ""

    def load(self, f=None, use_latest=True, ignore=False, flownet=False):
        if self.has_checkpoint() and use_latest:
            f = self.get_checkpoint_file()
        if not f:
            self.logger.info("No checkpoint found. Initializing model from scratch")
            return {}
        self.logger.info("Loading checkpoint from %s", f)
        checkpoint = self._load_file(f)
        load_state_dict(self.model, checkpoint.pop("model"), flownet=flownet)
        for k in ("optimizer", "scheduler"):              # inference build: training state is dropped, not restored
            checkpoint.pop(k, None)
        return checkpoint

    def load_flownet(self, f=None):
        self.logger.info("Loading flownet from %s", f)
        load_state_dict(self.model, torch.load(f, map_location="cpu")["state_dict"], flownet=True)

    def _load_file(self, f):
        return torch.load(f, map_location=torch.device("cpu"))


class DetectronCheckpointer(Checkpointer):
    def __init__(self, cfg, model, optimizer=None, scheduler=None, save_dir="", save_to_disk=None, logger=None):
        super().__init__(model, optimizer, scheduler, save_dir, save_to_disk, logger)
        self.cfg = cfg.clone()

    def _load_file(self, f):
        if f.startswith("catalog://") or f.startswith("http"):
            raise NotImplementedError("mega_core (B200 build): resolve %s to a local .pth file first (catalog / URL lookup "
                                      "is part of the reference's training-side tooling)" % f)
        if f.endswith(".pkl"):
            from .c2_model_loading import load_c2_format
            return load_c2_format(self.cfg, f)
        loaded = super()._load_file(f)
        return loaded if "model" in loaded else dict(model=loaded)
