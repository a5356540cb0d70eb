"""Shared helpers for the tests (numpy <-> split-fp16 NHWC, ctypes call wrappers)."""
import ctypes as C
import importlib

import numpy as np

pkg = importlib.import_module("deepcut-cnn_b200")
libdc = importlib.import_module("deepcut-cnn_b200.libdc")
gen_prototxt = importlib.import_module("deepcut-cnn_b200.gen_prototxt")
synth = importlib.import_module("deepcut-cnn_b200.synth")
ptx = importlib.import_module("deepcut-cnn_b200.prototxt")


def np_split(x_nchw):
    """fp32 NCHW -> (hi, lo) fp16 NHWC stacked as [2,N,H,W,C] (round-to-nearest, like split_f16)."""
    x = np.ascontiguousarray(np.asarray(x_nchw, np.float32).transpose(0, 2, 3, 1))
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return np.stack([hi, lo])


def np_join(s):
    """[2,N,H,W,C] fp16 -> fp32 NCHW."""
    return np.ascontiguousarray((s[0].astype(np.float32) + s[1].astype(np.float32)).transpose(0, 3, 1, 2))


def ptr(a):
    """numpy array or torch tensor -> void*"""
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


def pack_conv(w):
    L = libdc.lib()
    co, ci, kh, kw = w.shape
    rows = L.dc_packed_rows(co)
    K = kh * kw * ci
    packed = np.zeros((2, rows, K), np.uint16)
    rs = np.zeros(rows, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    libdc.check(L.dc_pack_conv_weight(ptr(w), co, ci, kh, kw, ptr(packed), ptr(rs)))
    return packed, rs


def pack_deconv(w):
    L = libdc.lib()
    ci, co, kh, kw = w.shape
    rows = L.dc_packed_rows(co * kh * kw)
    packed = np.zeros((2, rows, ci), np.uint16)
    rs = np.zeros(rows, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    libdc.check(L.dc_pack_deconv_weight(ptr(w), ci, co, kh, kw, ptr(packed), ptr(rs)))
    return packed, rs


def fold_bn(bn, sc, eps=1e-5):
    L = libdc.lib()
    c = bn[0].shape[0]
    a = np.zeros(c, np.float32)
    b = np.zeros(c, np.float32)
    g = np.ascontiguousarray(sc[0], np.float32) if sc else None
    be = np.ascontiguousarray(sc[1], np.float32) if sc and len(sc) > 1 else None
    libdc.check(L.dc_fold_bn_scale(ptr(np.ascontiguousarray(bn[0])), ptr(np.ascontiguousarray(bn[1])),
                                   float(bn[2][0]), eps, ptr(g) if g is not None else None,
                                   ptr(be) if be is not None else None, c, ptr(a), ptr(b)))
    return a, b


def caffe_module():
    """The pycaffe-compatible shim over libcaffe_b200.so."""
    import os
    import sys
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deepcut-cnn_b200", "python")
    if p not in sys.path:
        sys.path.insert(0, p)
    import caffe
    return caffe


def write_prototxt(tmpdir, **kw):
    import os
    path = os.path.join(str(tmpdir), "net_%s.prototxt" % "_".join(str(v) for v in kw.values()).replace(" ", "").replace(",", "-").strip("()"))
    gen_prototxt.write(path, **kw)
    return path
