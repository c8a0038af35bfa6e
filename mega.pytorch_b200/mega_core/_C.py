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


def sigmoid_focalloss_forward(logits, targets, num_classes, gamma, alpha):
    """(csrc/SigmoidFocalLoss.h:10-15) logits [N,C], int32 targets [N] -> losses [N,C]"""
    _cuda_only("sigmoid_focalloss_forward", logits, targets)
    logits = logits.contiguous().float()
    targets = targets.contiguous().to(torch.int32)
    out = torch.empty_like(logits)
    _lib.check(_lib.lib.mega_sigmoid_focalloss_forward(_lib.ptr(logits), _lib.ptr(targets), logits.shape[0],
                                                       int(num_classes), float(gamma), float(alpha), _lib.ptr(out),
                                                       _lib.stream_ptr()), "mega_sigmoid_focalloss_forward")
    return out


def sigmoid_focalloss_backward(logits, targets, d_losses, num_classes, gamma, alpha):
    """(csrc/SigmoidFocalLoss.h:26-32)"""
    _cuda_only("sigmoid_focalloss_backward", logits, targets, d_losses)
    logits = logits.contiguous().float()
    targets = targets.contiguous().to(torch.int32)
    d_losses = d_losses.contiguous().float()
    out = torch.empty_like(logits)
    _lib.check(_lib.lib.mega_sigmoid_focalloss_backward(_lib.ptr(logits), _lib.ptr(targets), _lib.ptr(d_losses),
                                                        logits.shape[0], int(num_classes), float(gamma), float(alpha),
                                                        _lib.ptr(out), _lib.stream_ptr()),
               "mega_sigmoid_focalloss_backward")
    return out


def _deform_conv(input, weight, bias, offset, mask, output, kh, kw, sh, sw, ph, pw, dh, dw, group, deformable_group):
    """bilinear im2col kernel + tcgen05 GEMM per group; writes `output` (NCHW) in place"""
    _cuda_only("deform_conv", input, weight, offset, output)
    if not (input.is_contiguous() and weight.is_contiguous()):
        raise RuntimeError("input / weight tensor has to be contiguous")     # deform_conv_cuda.cu:504-505
    b, c, h, w = input.shape
    cout, cpg = weight.shape[0], weight.shape[1]
    if tuple(weight.shape[2:]) != (kh, kw):
        raise RuntimeError("Input shape and kernel shape wont match: (%d x %d vs %d x %d)." % (kh, kw, weight.shape[2], weight.shape[3]))
    if c != cpg * group:
        raise RuntimeError("Input shape and kernel channels wont match: (%d vs %d)." % (c, cpg * group))
    ho = (h + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    wo = (w + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    kg = cpg * kh * kw
    cog = cout // group
    if kg % 4 or cog % 4:
        raise RuntimeError("deform_conv (B200 build): per-group C*kh*kw and Cout must be multiples of 4")
    kpad = c * kh * kw
    cols = torch.empty(b, ho * wo, kpad, device=input.device)
    _lib.check(_lib.lib.mega_deform_im2col(_lib.ptr(input.float()), _lib.ptr(offset.contiguous().float()),
                                           _lib.ptr(mask.contiguous().float()) if mask is not None else None, b, c, h,
                                           w, kh, kw, ph, pw, sh, sw, dh, dw, deformable_group, kpad, _lib.ptr(cols),
                                           _lib.stream_ptr()), "mega_deform_im2col")
    nhwc = torch.empty(b, 1, ho * wo, cout, device=input.device)
    w2 = weight.float().reshape(cout, kg)
    for g in range(group):
        a = cols.view(b, 1, ho * wo, kpad)[..., g * kg:(g + 1) * kg]
        wg = w2[g * cog:(g + 1) * cog].contiguous().view(1, cog, kg)
        bg = bias[g * cog:(g + 1) * cog].contiguous().float() if bias is not None else None
        ops.conv_gemm(a, wg, nhwc[..., g * cog:(g + 1) * cog], bias=bg, tile=(1, 128))
    out = output.view(b, cout, ho * wo)
    ops.transpose_2d(nhwc.view(b, ho * wo, cout), out, b, ho * wo, cout)
    return ho, wo


def deform_conv_forward(input, weight, offset, output, columns, ones, kW, kH, dW, dH, padW, padH, dilationW, dilationH,
                        group, deformable_group, im2col_step):
    """(csrc/deform_conv.h:11-28) v1; `output` [B,Cout,Ho,Wo] is written in place, returns 1"""
    _deform_conv(input, weight, None, offset, None, output, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group,
                 deformable_group)
    return 1


def modulated_deform_conv_forward(input, weight, bias, ones, offset, mask, output, columns, kernel_h, kernel_w,
                                  stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group, deformable_group,
                                  with_bias):
    """(csrc/deform_conv.h:115-150) v2; `output` is written in place"""
    _deform_conv(input, weight, bias if with_bias else None, offset, mask, output, kernel_h, kernel_w, stride_h, stride_w,
                 pad_h, pad_w, dilation_h, dilation_w, group, deformable_group)


def deform_psroi_pooling_forward(input, bbox, trans, out, top_count, no_trans, spatial_scale, output_dim, group_size,
                                 pooled_size, part_size, sample_per_part, trans_std):
    """(csrc/deform_pool.h:11-37) writes `out` [K,output_dim,ps,ps] and `top_count` in place"""
    _cuda_only("deform_psroi_pooling_forward", input, bbox, out, top_count)
    if not input.is_contiguous():
        raise RuntimeError("input must be contiguous")
    n, c, h, w = input.shape
    num_classes = 1 if no_trans else trans.shape[1] // 2
    tr = trans.contiguous().float() if not no_trans else None
    _lib.check(_lib.lib.mega_deform_psroi_pooling_forward(
        _lib.ptr(input.float()), _lib.ptr(bbox.contiguous().float()), _lib.ptr(tr), bbox.shape[0], c, h, w,
        int(bool(no_trans)), float(spatial_scale), int(output_dim), int(group_size), int(pooled_size), int(part_size),
        int(sample_per_part), float(trans_std), num_classes, _lib.ptr(out), _lib.ptr(top_count), _lib.stream_ptr()),
        "mega_deform_psroi_pooling_forward")


def _not_yet(name):
    def f(*a, **k):
        raise NotImplementedError("mega_core._C.%s: training-side / non-VID op, outside the inference hot path "
                                  "(SURVEY.md section 8f)" % name)
    f.__name__ = name
    return f


roi_align_backward = _not_yet("roi_align_backward")
roi_pool_forward = _not_yet("roi_pool_forward")
roi_pool_backward = _not_yet("roi_pool_backward")
deform_conv_backward_input = _not_yet("deform_conv_backward_input")
deform_conv_backward_parameters = _not_yet("deform_conv_backward_parameters")
modulated_deform_conv_backward = _not_yet("modulated_deform_conv_backward")
deform_psroi_pooling_backward = _not_yet("deform_psroi_pooling_backward")
