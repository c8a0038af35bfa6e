"""Layer wrappers with the reference's names (mega_core/layers/__init__.py:4-46): nms and the parameter holders here,
the differentiable `_C` ops (ROIAlign / ROIPool / focal loss / deformable convolution and pooling) in train_ops.py."""
import torch
from torch import nn

from .. import _C


def nms(boxes, scores, threshold):
    """layers/nms.py:8 of the reference (`amp.float_function(_C.nms)`): always fp32"""
    return _C.nms(boxes.float(), scores.float(), threshold)


from .train_ops import (ROIAlign, ROIPool, SigmoidFocalLoss, DeformConv, ModulatedDeformConv,  # noqa: E402,F401
                        ModulatedDeformConvPack, DeformRoIPooling, DeformRoIPoolingPack,
                        ModulatedDeformRoIPoolingPack, deform_conv, modulated_deform_conv, deform_roi_pooling,
                        roi_align, roi_pool, sigmoid_focal_loss_cuda, smooth_l1_loss)


class FrozenBatchNorm2d(nn.Module):
    """buffer holder with the reference's names (layers/batch_norm.py:5-31); the arithmetic
    (scale = weight * rsqrt(running_var), no eps) is folded into the conv epilogue on the device."""

    def __init__(self, n):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))


Conv2d = nn.Conv2d   # the reference's empty-batch-safe subclass (layers/misc.py:30-43) is a parameter holder here


__all__ = ["nms", "roi_align", "ROIAlign", "roi_pool", "ROIPool", "smooth_l1_loss", "Conv2d", "FrozenBatchNorm2d",
           "SigmoidFocalLoss", "deform_conv", "modulated_deform_conv", "DeformConv", "ModulatedDeformConv",
           "ModulatedDeformConvPack", "deform_roi_pooling", "DeformRoIPooling", "DeformRoIPoolingPack",
           "ModulatedDeformRoIPoolingPack"]
