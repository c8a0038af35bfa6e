"""train_ops_oracle -- CPU oracles for the training-side half of `mega_core._C`.

TEST INFRASTRUCTURE ONLY: nothing under `mega.pytorch_b200/` imports this module.

The reference implements these ops only in CUDA (csrc/ROIAlign.h:44, ROIPool.h:21, deform_conv.h:41, deform_pool.h:37
raise "Not implemented on the CPU") and its tests do not cover them, so there is no golden vector and nothing to compile
into oracle/_ref: PARITY UNPINNED by the reference itself. Each oracle below is therefore a differentiable PyTorch
restatement of the reference's FORWARD kernel (cited per function); the backward oracle is autograd of that restatement,
which is mathematically independent of the hand-derived gradient formulas in csrc/train_ops.cuh. The forward
restatements are anchored where an anchor exists (tests/test_train_ops_cpu.py):
  * roi_align      == oracle_roi_align_fwd (bit-pinned to the reference's ROIAlign_cpu.cpp compiled verbatim),
                      and == torchvision.ops.roi_align(aligned=False);
  * roi_pool       == torchvision.ops.roi_pool (same Caffe2 lineage as csrc/cuda/ROIPool_cuda.cu);
  * deform_conv2d  == torchvision.ops.deform_conv2d, forward and all five gradients (same mmdetection lineage as
                      csrc/cuda/deform_conv_kernel_cuda.cu);
  * deform_psroi_pool == mega_oracle.deform_psroi_pool (the plain-Python restatement already used for the forward op).
Sizes: small cases only (Python loops over rois / bins / samples).
"""
import math

import torch


# --------------------------------------------------------------------------------------------------- ROIAlign
def _bilinear_terms(height, width, y, x):
    """bilinear_interpolate of csrc/cuda/ROIAlign_cuda.cu:16-62 -> [(y, x, weight)] (empty outside the map)"""
    if y < -1.0 or y > height or x < -1.0 or x > width:
        return []
    y = max(y, 0.0)
    x = max(x, 0.0)
    y_low, x_low = int(y), int(x)
    if y_low >= height - 1:
        y_high = y_low = height - 1
        y = float(y_low)
    else:
        y_high = y_low + 1
    if x_low >= width - 1:
        x_high = x_low = width - 1
        x = float(x_low)
    else:
        x_high = x_low + 1
    ly, lx = y - y_low, x - x_low
    hy, hx = 1.0 - ly, 1.0 - lx
    return [(y_low, x_low, hy * hx), (y_low, x_high, hy * lx), (y_high, x_low, ly * hx), (y_high, x_high, ly * lx)]


def roi_align(feat, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio):
    """differentiable ROIAlign forward, NCHW in, [K,C,ph,pw] out (RoIAlignForward, ROIAlign_cuda.cu:64-122).
    Geometry in float32 like the kernel; the interpolation runs in feat.dtype."""
    f = torch.float32
    k = rois.shape[0]
    n, c, h, w = feat.shape
    rows = []
    for i in range(k):
        r = rois[i].to(f)
        b = int(r[0])
        sc = torch.tensor(spatial_scale, dtype=f)
        sw, sh, ew, eh = (r[1] * sc).item(), (r[2] * sc).item(), (r[3] * sc).item(), (r[4] * sc).item()
        t = lambda v: torch.tensor(v, dtype=f)           # noqa: E731  (float32 scalar arithmetic)
        rw = max((t(ew) - t(sw)).item(), 1.0)
        rh = max((t(eh) - t(sh)).item(), 1.0)
        bh, bw = (t(rh) / t(float(pooled_h))).item(), (t(rw) / t(float(pooled_w))).item()
        gh = sampling_ratio if sampling_ratio > 0 else int(math.ceil((t(rh) / pooled_h).item()))
        gw = sampling_ratio if sampling_ratio > 0 else int(math.ceil((t(rw) / pooled_w).item()))
        bins = []
        for ph in range(pooled_h):
            for pw in range(pooled_w):
                acc = feat.new_zeros(c)
                for iy in range(gh):
                    y = (t(sh) + t(float(ph)) * t(bh) + t(iy + 0.5) * t(bh) / t(float(gh))).item()
                    for ix in range(gw):
                        x = (t(sw) + t(float(pw)) * t(bw) + t(ix + 0.5) * t(bw) / t(float(gw))).item()
                        for (yy, xx, wt) in _bilinear_terms(h, w, y, x):
                            acc = acc + wt * feat[b, :, yy, xx]
                bins.append(acc / float(gh * gw))
        rows.append(torch.stack(bins, 1).reshape(c, pooled_h, pooled_w))
    if not rows:
        return feat.new_zeros(0, c, pooled_h, pooled_w)
    return torch.stack(rows, 0)


def roi_align_backward(grad, rois, spatial_scale, pooled_h, pooled_w, batch, channels, height, width, sampling_ratio):
    """`_C.roi_align_backward` (csrc/ROIAlign.h:27-45): d/d(input) of sum(roi_align(input) * grad)"""
    x = torch.zeros(batch, channels, height, width, dtype=torch.float64, requires_grad=True)
    out = roi_align(x, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio)
    if out.numel() == 0:
        return torch.zeros(batch, channels, height, width)
    (out * grad.double()).sum().backward()
    return x.grad.float()


# ---------------------------------------------------------------------------------------------------- ROIPool
def _c_round(v):
    """C round(): half away from zero"""
    return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)


def roi_pool(feat, rois, spatial_scale, pooled_h, pooled_w):
    """differentiable ROIPool forward -> (out [K,C,ph,pw], argmax int32) (RoIPoolFForward, ROIPool_cuda.cu:16-77)"""
    f = torch.float32
    k = rois.shape[0]
    n, c, h, w = feat.shape
    out_rows, arg_rows = [], []
    for i in range(k):
        r = rois[i].to(f)
        b = int(r[0])
        sc = torch.tensor(spatial_scale, dtype=f)
        sw, sh = _c_round((r[1] * sc).item()), _c_round((r[2] * sc).item())
        ew, eh = _c_round((r[3] * sc).item()), _c_round((r[4] * sc).item())
        rw, rh = max(ew - sw + 1, 1), max(eh - sh + 1, 1)
        bh = (torch.tensor(float(rh), dtype=f) / torch.tensor(float(pooled_h), dtype=f)).item()
        bw = (torch.tensor(float(rw), dtype=f) / torch.tensor(float(pooled_w), dtype=f)).item()
        vals, args = [], []
        for ph in range(pooled_h):
            for pw in range(pooled_w):
                f32 = lambda v: torch.tensor(v, dtype=f).item()   # noqa: E731
                hs = int(math.floor(f32(f32(float(ph)) * bh)))
                ws = int(math.floor(f32(f32(float(pw)) * bw)))
                he = int(math.ceil(f32(f32(float(ph + 1)) * bh)))
                we = int(math.ceil(f32(f32(float(pw + 1)) * bw)))
                hs, he = min(max(hs + sh, 0), h), min(max(he + sh, 0), h)
                ws, we = min(max(ws + sw, 0), w), min(max(we + sw, 0), w)
                if he <= hs or we <= ws:
                    vals.append(feat.new_zeros(c))
                    args.append(torch.full((c,), -1, dtype=torch.int32))
                    continue
                win = feat[b, :, hs:he, ws:we].reshape(c, -1)
                v, idx = win.max(dim=1)
                yy = idx // (we - ws) + hs
                xx = idx % (we - ws) + ws
                vals.append(v)
                args.append((yy * w + xx).to(torch.int32))
        out_rows.append(torch.stack(vals, 1).reshape(c, pooled_h, pooled_w))
        arg_rows.append(torch.stack(args, 1).reshape(c, pooled_h, pooled_w))
    if not out_rows:
        return feat.new_zeros(0, c, pooled_h, pooled_w), torch.zeros(0, c, pooled_h, pooled_w, dtype=torch.int32)
    return torch.stack(out_rows, 0), torch.stack(arg_rows, 0)


def roi_pool_backward(grad, feat, rois, spatial_scale, pooled_h, pooled_w):
    """`_C.roi_pool_backward` (csrc/ROIPool.h:26-47): autograd of roi_pool (gradient goes to each bin's arg-max)"""
    x = feat.detach().double().requires_grad_(True)
    out, _ = roi_pool(x, rois, spatial_scale, pooled_h, pooled_w)
    if out.numel() == 0:
        return torch.zeros_like(feat)
    (out * grad.double()).sum().backward()
    return x.grad.float()


# ------------------------------------------------------------------------------- deformable convolution v1 / v2
def deform_conv2d(x, offset, mask, weight, bias, stride, padding, dilation, groups, deformable_groups):
    """differentiable (modulated) deformable convolution, vectorised.
    Sampling: deformable_im2col_bilinear + the `h_im > -1 && w_im > -1 && h_im < height && w_im < width` gate of
    deformable_im2col_gpu_kernel / modulated_deformable_im2col_gpu_kernel (deform_conv_kernel_cuda.cu:92-120, :197-250,
    :475-504, :578-640); contraction: the per-group addmm_ of deform_conv_cuda.cu:228-243, :545-567.
    x [B,C,H,W]; offset [B, dg*2*kh*kw, Ho, Wo] ((dh, dw) interleaved per tap); mask [B, dg*kh*kw, Ho, Wo] or None;
    weight [Cout, C/groups, kh, kw]; stride / padding / dilation: (h, w) pairs."""
    b, c, h, w = x.shape
    cout, cpg_w, kh, kw = weight.shape
    sh, sw = stride
    ph, pw = padding
    dh, dw = dilation
    ho = (h + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    wo = (w + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    dg = deformable_groups
    cpg = c // dg
    taps = kh * kw
    dt = x.dtype
    base_h = (torch.arange(ho, dtype=dt) * sh - ph).view(1, 1, 1, ho, 1)
    base_w = (torch.arange(wo, dtype=dt) * sw - pw).view(1, 1, 1, 1, wo)
    ki = (torch.arange(kh, dtype=dt) * dh).repeat_interleave(kw).view(1, 1, taps, 1, 1)
    kj = (torch.arange(kw, dtype=dt) * dw).repeat(kh).view(1, 1, taps, 1, 1)
    off = offset.view(b, dg, taps, 2, ho, wo)
    hh = base_h + ki + off[:, :, :, 0]                       # [B, dg, taps, Ho, Wo]
    ww = base_w + kj + off[:, :, :, 1]
    inside = (hh > -1) & (ww > -1) & (hh < h) & (ww < w)
    h_low, w_low = torch.floor(hh), torch.floor(ww)
    lh, lw = hh - h_low, ww - w_low
    h_low, w_low = h_low.long(), w_low.long()
    xg = x.view(b, dg, cpg, h * w)

    def corner(yy, xx, wt):
        ok = inside & (yy >= 0) & (yy <= h - 1) & (xx >= 0) & (xx <= w - 1)
        idx = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)).view(b, dg, 1, -1).expand(b, dg, cpg, -1)
        v = torch.gather(xg, 3, idx).view(b, dg, cpg, taps, ho, wo)
        return v * (wt * ok.to(dt)).unsqueeze(2)

    val = (corner(h_low, w_low, (1 - lh) * (1 - lw)) + corner(h_low, w_low + 1, (1 - lh) * lw) +
           corner(h_low + 1, w_low, lh * (1 - lw)) + corner(h_low + 1, w_low + 1, lh * lw))
    if mask is not None:
        val = val * mask.view(b, dg, 1, taps, ho, wo)
    cols = val.reshape(b, c * taps, ho * wo)                  # k = c*taps + tap
    cog = cout // groups
    kg = (c // groups) * taps
    outs = []
    for g in range(groups):
        wg = weight[g * cog:(g + 1) * cog].reshape(cog, kg)
        outs.append(torch.einsum("ok,bkp->bop", wg, cols[:, g * kg:(g + 1) * kg]))
    out = torch.cat(outs, 1).view(b, cout, ho, wo)
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)
    return out


def deform_conv2d_grads(x, offset, mask, weight, bias, grad_out, stride, padding, dilation, groups, deformable_groups):
    """all gradients of deform_conv2d by autograd, computed in float64 and returned as float32:
    dict(input, offset, mask, weight, bias) -- what `_C.deform_conv_backward_input` / `_backward_parameters` /
    `modulated_deform_conv_backward` return (deform_conv.h:45-113, :152-190)."""
    d = torch.float64
    leaves = {"input": x, "offset": offset, "mask": mask, "weight": weight, "bias": bias}
    leaves = {k: (v.detach().to(d).requires_grad_(True) if v is not None else None) for k, v in leaves.items()}
    out = deform_conv2d(leaves["input"], leaves["offset"], leaves["mask"], leaves["weight"], leaves["bias"], stride,
                        padding, dilation, groups, deformable_groups)
    (out * grad_out.to(d)).sum().backward()
    return {k: (v.grad.float() if v is not None else None) for k, v in leaves.items()}


# ------------------------------------------------------------------------------- deformable PSROI pooling
def deform_psroi_pool(data, rois, trans, no_trans, spatial_scale, output_dim, group_size, pooled_size, part_size,
                      sample_per_part, trans_std):
    """differentiable restatement of DeformablePSROIPoolForwardKernel (csrc/cuda/deform_pool_kernel_cuda.cu:31-142)
    -> (out [K, output_dim, ps, ps], top_count). ROI geometry in float32 like the kernel; the translation, the sample
    positions and the interpolation in data.dtype so that autograd reaches `trans`."""
    f = torch.float32
    dt = data.dtype
    n_rois = rois.shape[0]
    _, channels, height, width = data.shape
    num_classes = 1 if no_trans else trans.shape[1] // 2
    cec = output_dim if no_trans else output_dim // num_classes
    out = []
    cnt = torch.zeros(n_rois, output_dim, pooled_size, pooled_size)
    t32 = lambda v: torch.tensor(v, dtype=f)                       # noqa: E731
    for n in range(n_rois):
        b = int(rois[n, 0])
        sc = t32(spatial_scale)
        sw = (t32(float(_c_round(float(rois[n, 1])))) * sc - 0.5).item()
        sh = (t32(float(_c_round(float(rois[n, 2])))) * sc - 0.5).item()
        ew = (t32(float(_c_round(float(rois[n, 3])) + 1.0)) * sc - 0.5).item()
        eh = (t32(float(_c_round(float(rois[n, 4])) + 1.0)) * sc - 0.5).item()
        rw = max((t32(ew) - t32(sw)).item(), t32(0.1).item())
        rh = max((t32(eh) - t32(sh)).item(), t32(0.1).item())
        bh, bw = (t32(rh) / pooled_size).item(), (t32(rw) / pooled_size).item()
        sbh, sbw = (t32(bh) / sample_per_part).item(), (t32(bw) / sample_per_part).item()
        for ctop in range(output_dim):
            cls = ctop // cec
            for ph in range(pooled_size):
                for pw in range(pooled_size):
                    part_h = int(math.floor((t32(float(ph)) / pooled_size * part_size).item()))
                    part_w = int(math.floor((t32(float(pw)) / pooled_size * part_size).item()))
                    if no_trans:
                        tx = ty = torch.zeros((), dtype=dt)
                    else:
                        tx = trans[n, cls * 2, part_h, part_w] * trans_std
                        ty = trans[n, cls * 2 + 1, part_h, part_w] * trans_std
                    wstart = (t32(float(pw)) * t32(bw) + t32(sw)).item() + tx * rw
                    hstart = (t32(float(ph)) * t32(bh) + t32(sh)).item() + ty * rh
                    gw = min(max(int(math.floor((t32(float(pw)) * group_size / pooled_size).item())), 0), group_size - 1)
                    gh = min(max(int(math.floor((t32(float(ph)) * group_size / pooled_size).item())), 0), group_size - 1)
                    c = (ctop * group_size + gh) * group_size + gw
                    s = torch.zeros((), dtype=dt)
                    k = 0
                    for ih in range(sample_per_part):
                        for iw in range(sample_per_part):
                            w_ = wstart + iw * sbw
                            h_ = hstart + ih * sbh
                            wv, hv = float(w_.detach()), float(h_.detach())
                            if wv < -0.5 or wv > width - 0.5 or hv < -0.5 or hv > height - 0.5:
                                continue
                            w_ = w_.clamp(0.0, width - 1.0)
                            h_ = h_.clamp(0.0, height - 1.0)
                            x1, x2 = int(math.floor(float(w_.detach()))), int(math.ceil(float(w_.detach())))
                            y1, y2 = int(math.floor(float(h_.detach()))), int(math.ceil(float(h_.detach())))
                            dx, dy = w_ - x1, h_ - y1
                            pl = data[b, c]
                            s = s + (1 - dx) * (1 - dy) * pl[y1, x1] + (1 - dx) * dy * pl[y2, x1] + \
                                dx * (1 - dy) * pl[y1, x2] + dx * dy * pl[y2, x2]
                            k += 1
                    out.append(s / k if k else torch.zeros((), dtype=dt))
                    cnt[n, ctop, ph, pw] = k
    out = torch.stack(out).view(n_rois, output_dim, pooled_size, pooled_size) if out else \
        torch.zeros(0, output_dim, pooled_size, pooled_size, dtype=dt)
    return out, cnt


def deform_psroi_pool_grads(data, rois, trans, out_grad, no_trans, spatial_scale, output_dim, group_size, pooled_size,
                            part_size, sample_per_part, trans_std):
    """(input_grad, trans_grad) of `_C.deform_psroi_pooling_backward` (csrc/deform_pool.h:41-69) by autograd, float64"""
    d = torch.float64
    x = data.detach().to(d).requires_grad_(True)
    t = trans.detach().to(d).requires_grad_(True) if not no_trans else None
    out, _ = deform_psroi_pool(x, rois, t, no_trans, spatial_scale, output_dim, group_size, pooled_size, part_size,
                               sample_per_part, trans_std)
    (out * out_grad.to(d)).sum().backward()
    return x.grad.float(), (t.grad.float() if t is not None and t.grad is not None else None)
