"""GPU-side harness for the kernel parity tests: runs single layers through the C ABI with torch
holding device memory (torch = plumbing only)."""
import ctypes as C

import numpy as np
import torch

import dcutil
from dcutil import libdc, ptr


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def init():
    libdc.check(libdc.lib().dc_init(torch.cuda.current_device()))


def conv_bn(x_nchw, w, a, b, pad=0, dil=1, relu=True, residual_nchw=None, f32_rows=False, stride=1, split_k_workspace=True):
    """Runs dc_conv_forward on fp32 NCHW numpy inputs; returns fp32 NCHW numpy (or fp32 rows)."""
    L = libdc.lib()
    n, ci, h, wd = x_nchw.shape
    co, _, kh, kw = w.shape
    packed, rs = dcutil.pack_conv(w)
    rows = packed.shape[1]
    scale = np.ones(rows, np.float32)
    shift = np.zeros(rows, np.float32)
    scale[:co] = a * rs[:co]
    shift[:co] = b
    ho = (h + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    wo = (wd + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    xs = dev(dcutil.np_split(x_nchw))
    wp, sc, sh = dev(packed), dev(scale), dev(shift)
    res = dev(dcutil.np_split(residual_nchw)) if residual_nchw is not None else None
    if f32_rows:
        out = torch.full((n * ho * wo, rows), float("nan"), dtype=torch.float32, device="cuda")
    else:
        out = torch.full((2, n, ho, wo, co), float("nan"), dtype=torch.float16, device="cuda")
    ws_bytes = L.dc_splitk_workspace_bytes()
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")       # split-K scratch (uninitialised on purpose)
    args = libdc.ConvArgs(x=xs.data_ptr(), n=n, h=h, w=wd, cin=ci, cout=co, kh=kh, kw=kw, pad=pad, dilation=dil,
                          w_packed=wp.data_ptr(), scale=sc.data_ptr(), shift=sh.data_ptr(),
                          residual=res.data_ptr() if res is not None else None, relu=int(relu),
                          out_f32_rows=int(f32_rows), ldc=rows, out=out.data_ptr(), stride=stride,
                          splitk_workspace=ws.data_ptr() if split_k_workspace else None, splitk_workspace_bytes=ws_bytes)
    libdc.check(L.dc_conv_forward(C.byref(args), stream_ptr()))
    torch.cuda.synchronize()
    if f32_rows:
        return out.cpu().numpy()
    return dcutil.np_join(out.cpu().numpy())


def conv_bn_subbatch(x_nchw, w, a, b, i0, cn, pad=0, dil=1, relu=True, residual_nchw=None, inplace=False):
    """dc_conv_forward over images [i0, i0 + cn) of full-batch tensors (sub-batch pointers + full-tensor plane strides, the
    chunked schedule's launch form).  The output tensor starts as NaN (or as the residual when `inplace`: the block output
    overwrites its shortcut); returns the WHOLE output tensor as split fp16 [2, N, Ho, Wo, Co] (numpy)."""
    L = libdc.lib()
    n, ci, h, wd = x_nchw.shape
    co, _, kh, kw = w.shape
    packed, rs = dcutil.pack_conv(w)
    rows = packed.shape[1]
    scale = np.ones(rows, np.float32)
    shift = np.zeros(rows, np.float32)
    scale[:co] = a * rs[:co]
    shift[:co] = b
    ho = h + 2 * pad - (dil * (kh - 1) + 1) + 1
    wo = wd + 2 * pad - (dil * (kw - 1) + 1) + 1
    xs = dev(dcutil.np_split(x_nchw))
    wp, sc, sh = dev(packed), dev(scale), dev(shift)
    res = dev(dcutil.np_split(residual_nchw)) if residual_nchw is not None else None
    out = res if inplace else torch.full((2, n, ho, wo, co), float("nan"), dtype=torch.float16, device="cuda")
    ws_bytes = L.dc_splitk_workspace_bytes()
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    x_img, o_img = h * wd * ci, ho * wo * co
    args = libdc.ConvArgs(x=xs.data_ptr() + 2 * i0 * x_img, n=cn, h=h, w=wd, cin=ci, cout=co, kh=kh, kw=kw, pad=pad, dilation=dil,
                          w_packed=wp.data_ptr(), scale=sc.data_ptr(), shift=sh.data_ptr(),
                          residual=(res.data_ptr() + 2 * i0 * o_img) if res is not None else None, relu=int(relu),
                          out_f32_rows=0, ldc=rows, out=out.data_ptr() + 2 * i0 * o_img, stride=1,
                          splitk_workspace=ws.data_ptr(), splitk_workspace_bytes=ws_bytes,
                          x_plane=n * x_img, out_plane=n * o_img, residual_plane=n * o_img if res is not None else 0)
    libdc.check(L.dc_conv_forward(C.byref(args), stream_ptr()))
    torch.cuda.synchronize()
    return out.cpu().numpy()
