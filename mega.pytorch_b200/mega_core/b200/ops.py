"""Python wrappers (pointer plumbing only) around the C-ABI kernels."""
import ctypes

import torch

from .. import _lib
from .._lib import ConvGemmDesc, check, lib, ptr, require_cuda, stream_ptr


def pick_tile(h, w):
    """128-pixel output tile (tile_h, tile_w) wasting the fewest pixels for an h x w map."""
    best = None
    for tw in (128, 64, 32, 16, 8, 4, 2, 1):
        th = 128 // tw
        tiles = -(-h // th) * -(-w // tw)
        key = (tiles, -tw)
        if best is None or key < best[0]:
            best = (key, th, tw)
    return best[1], best[2]


def pick_block_n(cout, m_tiles, batch=1):
    """Largest N tile that still gives every SM a CTA (148 SMs); small layers favour more CTAs."""
    for bn in (256, 128, 64):
        if cout >= bn and m_tiles * -(-cout // bn) * batch >= 148:
            return bn
    for bn in (32, 64, 128, 256):
        if cout <= bn:
            return bn
    return 64


def conv_gemm(a, w, out, *, taps=(1, 1), dil=1, pad=0, scale=None, bias=None, residual=None,
              relu=False, tile=None, block_n=None, cout=None, k=None, batch=1, a_c_off=0,
              a_n_off=0, b_k_off=0, b_n_off=0, out_z_off=0, res_z_off=0, splits=1, partial=None,
              out_hw=None):
    """out[n,h,w,:] = act(scale * conv(a, w) + bias + residual)   (TF32 tensor cores)

    a   : [N,H,W,C] fp32 view (innermost stride 1; other strides multiples of 4 floats)
    w   : [taps, rows, K] fp32 (K contiguous)
    out : [N,Ho,Wo,>=cout] fp32 view (innermost stride 1)
    """
    require_cuda(a, w, out, scale, bias, residual, partial)
    assert a.dtype == torch.float32 and w.dtype == torch.float32 and out.dtype == torch.float32
    assert a.dim() == 4 and w.dim() == 3 and out.dim() == 4
    assert a.stride(3) == 1 and w.stride(2) == 1 and out.stride(3) == 1
    n, h, wd, c = a.shape
    t, rows, kk = w.shape
    assert t == taps[0] * taps[1]
    on, oh, ow, oc = out.shape
    if out_hw is not None:
        oh, ow = out_hw
    assert out.stride(1) == out.stride(2) * out.shape[2] or out.shape[1] == 1
    assert out.stride(0) == out.stride(2) * out.shape[2] * out.shape[1] or out.shape[0] == 1
    d = ConvGemmDesc()
    d.a = ptr(a)
    d.a_n, d.a_h, d.a_w, d.a_c = n, h, wd, c
    d.a_stride_w, d.a_stride_h, d.a_stride_n = a.stride(2), a.stride(1), a.stride(0)
    d.b = ptr(w)
    d.b_n, d.b_k = rows, kk
    d.b_stride_n, d.b_stride_tap = w.stride(1), w.stride(0)
    d.taps_r, d.taps_s, d.dil, d.pad = taps[0], taps[1], dil, pad
    d.k_per_tap = k if k is not None else kk
    d.out = ptr(out)
    d.out_ld = out.stride(2)
    d.n_img, d.out_h, d.out_w = on, oh, ow
    d.cout = cout if cout is not None else rows
    d.scale, d.bias, d.residual = ptr(scale), ptr(bias), ptr(residual)
    d.res_ld = residual.stride(-2) if residual is not None else 0
    d.relu = 1 if relu else 0
    th, tw = tile if tile is not None else pick_tile(oh, ow)
    d.tile_h, d.tile_w = th, tw
    m_tiles = on * (-(-oh // th)) * (-(-ow // tw))
    d.block_n = block_n if block_n is not None else pick_block_n(d.cout, m_tiles, batch)
    d.batch = batch
    d.a_c_off, d.a_n_off, d.b_k_off, d.b_n_off = a_c_off, a_n_off, b_k_off, b_n_off
    d.out_z_off, d.res_z_off = out_z_off, res_z_off
    d.splits = splits
    d.partial = ptr(partial)
    check(lib.mega_conv_gemm_tf32(ctypes.byref(d), stream_ptr()), "mega_conv_gemm_tf32")
    return out


def linear(x, w, out, *, bias=None, relu=False, residual=None, splits=1, partial=None, block_n=None):
    """out[m,:] = act(x[m,:] @ w.T + bias + residual[m,:]); x [M,K], w [N,K], out [M,N]."""
    m, kdim = x.shape
    nrows = w.shape[0]
    a4 = x.as_strided((1, 1, m, kdim), (x.stride(0) * m, x.stride(0) * m, x.stride(0), 1))
    o4 = out.as_strided((1, 1, m, out.shape[1]), (out.stride(0) * m, out.stride(0) * m, out.stride(0), 1))
    r4 = None
    if residual is not None:
        r4 = residual.as_strided((1, 1, m, residual.shape[1]),
                                 (residual.stride(0) * m, residual.stride(0) * m, residual.stride(0), 1))
    w3 = w.as_strided((1, nrows, kdim), (w.stride(0) * nrows, w.stride(0), 1))
    return conv_gemm(a4, w3, o4, bias=bias, relu=relu, residual=r4, tile=(1, 128), cout=nrows,
                     splits=splits, partial=partial, block_n=block_n)
