"""`mega_core._C` -- the reference's native-op module (csrc/vision.cpp:9-25), served by the sm_100a
kernels of libmega_b200.so. Signatures are positional and identical to the reference's pybind
functions; device tensors only: there is no CPU implementation behind these names."""
import torch

from . import _lib
from .b200 import ops


def _cuda_only(name, *tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("%s: Not implemented on the CPU (mega_core._C is the B200 build; "
                               "the CPU path of the reference lives only in the test oracle)" % name)


def nms(dets, scores, threshold):
    """nms(Tensor dets[n,4], Tensor scores[n], float thr) -> LongTensor (csrc/nms.h:10-28).
    Kept original indices in ascending order, on the input's device (nms.cu:127-130); an empty
    input returns an empty long tensor."""
    if dets.numel() == 0:
        return torch.empty(0, dtype=torch.int64, device=dets.device)
    _cuda_only("nms", dets, scores)
    keep, count = ops.nms_device(dets, scores, float(threshold))
    return keep[: int(count.item())]        # the one host read the dynamic output size requires


def roi_align_forward(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio):
    """(csrc/ROIAlign.h:11-25) NCHW fp32 input, rois [K,5] = (batch, x1, y1, x2, y2) -> [K,C,ph,pw]"""
    _cuda_only("roi_align_forward", input, rois)
    return ops.roi_align_nchw(input, rois, float(spatial_scale), int(pooled_height), int(pooled_width),
                              int(sampling_ratio))


def _not_yet(name):
    def f(*a, **k):
        raise NotImplementedError("mega_core._C.%s: training-side / non-VID op, outside the inference hot path "
                                  "(SURVEY.md section 8f)" % name)
    f.__name__ = name
    return f


roi_align_backward = _not_yet("roi_align_backward")
roi_pool_forward = _not_yet("roi_pool_forward")
roi_pool_backward = _not_yet("roi_pool_backward")
sigmoid_focalloss_forward = _not_yet("sigmoid_focalloss_forward")
sigmoid_focalloss_backward = _not_yet("sigmoid_focalloss_backward")
deform_conv_forward = _not_yet("deform_conv_forward")
deform_conv_backward_input = _not_yet("deform_conv_backward_input")
deform_conv_backward_parameters = _not_yet("deform_conv_backward_parameters")
modulated_deform_conv_forward = _not_yet("modulated_deform_conv_forward")
modulated_deform_conv_backward = _not_yet("modulated_deform_conv_backward")
deform_psroi_pooling_forward = _not_yet("deform_psroi_pooling_forward")
deform_psroi_pooling_backward = _not_yet("deform_psroi_pooling_backward")
