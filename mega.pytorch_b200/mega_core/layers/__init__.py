"""Layer wrappers with the reference's names (mega_core/layers/__init__.py:4-46), inference side."""
import torch
from torch import nn

from .. import _C


def nms(boxes, scores, threshold):
    """layers/nms.py:8 of the reference (`amp.float_function(_C.nms)`): always fp32"""
    return _C.nms(boxes.float(), scores.float(), threshold)


def roi_align(input, rois, output_size, spatial_scale, sampling_ratio):
    oh, ow = (output_size, output_size) if isinstance(output_size, int) else output_size
    return _C.roi_align_forward(input.float(), rois.float(), spatial_scale, oh, ow, sampling_ratio)


class ROIAlign(nn.Module):
    """layers/roi_align.py:47-60"""

    def __init__(self, output_size, spatial_scale, sampling_ratio):
        super().__init__()
        self.output_size, self.spatial_scale, self.sampling_ratio = output_size, spatial_scale, sampling_ratio

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio)

    def __repr__(self):
        return "ROIAlign(output_size=%s, spatial_scale=%s, sampling_ratio=%s)" % (
            self.output_size, self.spatial_scale, self.sampling_ratio)


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
