"""CPU restatement of the demo's pre-processing, python/pose/estimate_pose.py:83-105 (TEST INFRASTRUCTURE ONLY).

The arithmetic lives in a third-party dependency that is not under /root/reference: ``scipy.misc.imresize(image,
factor, interp='bilinear')`` (removed from SciPy 1.3; estimate_pose.py:98) is ``PIL.Image.resize(size, BILINEAR)`` on the
uint8 image with ``size = (int(W * factor), int(H * factor))``.  No version is pinned by the reference; this
restates Pillow's published 8-bit resampling algorithm (src/libImaging/Resample.c: precompute_coeffs,
normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc -- separable triangle filter whose support
widens when shrinking, 22-bit fixed-point coefficients, uint8 rounding after EACH pass, horizontal pass first) and
tests/test_preprocess_cpu.py pins it bit-exactly against the Pillow installed in this image (12.2.0).
"""
import math

import numpy as np

MEAN = np.array([104.0, 117.0, 123.0])       # estimate_pose.py:25 (applied to the array's channel order: BGR in the demo)
STRIDE = 8.0                                 # estimate_pose.py:31
PAD = 64                                     # estimate_pose.py:90
PRECISION_BITS = 32 - 8 - 2                  # Resample.c


def resample_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle, support 1) filter over the
    full source range.  -> (ksize, bounds int32 [out, 2] = (first source index, count), kk int32 [out, ksize])."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.zeros(xmax, np.float64)
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
        ww = 0.0                                # accumulated left to right, as Resample.c does
        for x in range(xmax):
            ww += w[x]
        if ww != 0.0:
            w = w / ww
        for x in range(xmax):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _resample_axis0(img, out_size):
    """One 8bpc pass along axis 0 of a uint8 array [in, ...]."""
    in_size = img.shape[0]
    _, bounds, kk = resample_coeffs(in_size, out_size)
    src = img.astype(np.int64)
    out = np.empty((out_size,) + img.shape[1:], np.uint8)
    for xx in range(out_size):
        xmin, xmax = bounds[xx]
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(xmax):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pillow_bilinear_resize_u8(img, out_w, out_h):
    """PIL.Image.fromarray(img).resize((out_w, out_h), BILINEAR) for uint8 [H, W, C]: horizontal pass, then vertical
    (ImagingResample: a pass is skipped when that dimension does not change)."""
    h, w = img.shape[:2]
    out = img
    if out_w != w:
        out = _resample_axis0(out.transpose(1, 0, 2), out_w).transpose(1, 0, 2)
    if out_h != h:
        out = _resample_axis0(out, out_h)
    return np.ascontiguousarray(out)


def net_input_size(h, w, scale):
    """estimate_pose.py:84-88: the net input is the scaled ORIGINAL size rounded up to the stride."""
    return (int(np.ceil(float(h) * scale / STRIDE) * STRIDE), int(np.ceil(float(w) * scale / STRIDE) * STRIDE))


def net_input_from_image(image, scale, mean=MEAN):
    """estimate_pose.py:83-105.  image uint8 [H, W, 3] (BGR in the demo) -> float32 [3, Hb, Wb] (the blob layout of
    _cnn_process_image, :225)."""
    h, w = image.shape[:2]
    bg_h, bg_w = net_input_size(h, w, scale)
    img = np.vstack((image, np.tile(image[-1:], (PAD, 1, 1))))          # :91-93 edge-replicate 64 rows below
    img = np.hstack((img, np.tile(img[:, -1:], (1, PAD, 1))))           # :94-96 and 64 columns right
    out_w, out_h = int(img.shape[1] * scale), int(img.shape[0] * scale)  # imresize: (array(im.size) * size).astype(int)
    img = pillow_bilinear_resize_u8(img, out_w, out_h).astype(np.float32) - mean      # :98-99
    net_input = np.zeros((bg_h, bg_w, 3), np.float32)                   # :101
    hh, ww = min(bg_h, img.shape[0]), min(bg_w, img.shape[1])
    net_input[:hh, :ww] = img[:hh, :ww]                                 # :102-105
    return np.ascontiguousarray(net_input.transpose(2, 0, 1))
