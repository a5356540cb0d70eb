"""Single-person pose estimation on the B200 path: the interface of the reference's python/pose/estimate_pose.py
(``estimate_pose(image, model_def, model_bin, scales=None)`` -> 5x14 pose), with both sides of the forward pass moved to
the device:

  * pre-processing (estimate_pose.py:83-105: edge-replicate pad 64, scipy.misc.imresize bilinear == Pillow's 8-bit
    resample, mean subtraction, crop to a multiple of the stride) = ``dc_preprocess_u8_forward`` straight into the `data`
    blob's device memory -- the host uploads the uint8 image (3 B/pixel) instead of the float net input (12 B/pixel);
  * read-out (:131-143 _pose_from_mats: per-joint arg-max + location refinement) = ``dc_pose_from_maps``; 280 bytes come back
    per scale instead of the `prob` + `loc_pred` maps.

One Net per input geometry is kept (reshaping rebuilds the fused plan), so a scale pyramid costs one plan per scale once.
The reference tiles inputs larger than 700 px (:160-221, a 2016 GPU-memory workaround whose seams change the result);
here every image runs whole.  No CPU fallback: without the CUDA library the import fails.
"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))                      # .../python  -> import caffe
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(_HERE))))    # repo root -> the package
import caffe as _caffe  # noqa: E402

_libdc = importlib.import_module("deepcut-cnn_b200.libdc")

_MEAN = np.array([104., 117., 123.], np.float32)       # estimate_pose.py:25
_LOCREF_SCALE_MUL = float(np.sqrt(53.))                 # :27
_STRIDE = 8.                                            # :31

_MODELS = {}          # (model_def, model_bin, out_h, out_w) -> Net
_PLANS = {}           # (h, w, scale) -> (plan handle, out_h, out_w, workspace bytes)


class _DeviceBuffer(object):
    def __init__(self):
        self.ptr, self.size = C.c_void_p(), 0

    def reserve(self, nbytes):
        if nbytes > self.size:
            L = _libdc.lib()
            if self.ptr:
                L.dc_free(self.ptr)
            self.ptr = C.c_void_p()
            _libdc.check(L.dc_malloc(C.byref(self.ptr), nbytes))
            self.size = nbytes
        return self.ptr


_IMAGE, _WORK, _POSE = _DeviceBuffer(), _DeviceBuffer(), _DeviceBuffer()


def _plan(h, w, scale):
    key = (h, w, float(scale))
    if key not in _PLANS:
        L = _libdc.lib()
        handle, oh, ow, ws = C.c_void_p(), C.c_int(), C.c_int(), C.c_size_t()
        _libdc.check(L.dc_preprocess_plan_create(h, w, float(scale), C.byref(handle)))
        _libdc.check(L.dc_preprocess_plan_info(handle, C.byref(oh), C.byref(ow), C.byref(ws)))
        _PLANS[key] = (handle, oh.value, ow.value, ws.value)
    return _PLANS[key]


def _model(model_def, model_bin, out_h, out_w, weights=None):
    key = (model_def, model_bin, out_h, out_w)
    if key not in _MODELS:
        net = _caffe.Net(model_def, model_bin, _caffe.TEST) if model_bin else _caffe.Net(model_def, _caffe.TEST)
        if weights is not None:
            net.set_params(weights)
        net.blobs['data'].reshape(1, 3, out_h, out_w)
        _MODELS[key] = net
    return _MODELS[key]


def preprocess_to_blob(image, scale, blob, stream):
    """image uint8 [H, W, 3] -> `blob` (reshaped to 1x3xHbxWb by the caller) on the device.  Returns (Hb, Wb)."""
    L = _libdc.lib()
    image = np.ascontiguousarray(image, np.uint8)
    h, w = image.shape[:2]
    plan, out_h, out_w, ws = _plan(h, w, scale)
    d_img = _IMAGE.reserve(image.nbytes)
    _libdc.check(L.dc_memcpy_async(d_img, image.ctypes.data_as(C.c_void_p), image.nbytes, 1, stream))
    d_ws = _WORK.reserve(ws) if ws else None
    _libdc.check(L.dc_preprocess_u8_forward(plan, d_img, _MEAN.ctypes.data_as(C.POINTER(C.c_float)), blob.mutable_gpu_data_ptr(), d_ws, stream))
    return out_h, out_w


def estimate_pose(image, model_def, model_bin, scales=None, weights=None):
    """Same contract as the reference's estimate_pose (estimate_pose.py:37-129): `image` uint8 HxWx3 in the channel
    order the model was trained on (BGR in the demo); returns the 5x14 array {x, y, confidence, offset x, offset y} of
    the scale whose weakest joint is most confident.  `weights` (layer -> arrays) stands in for `model_bin` when no
    trained .caffemodel exists."""
    if scales is None:
        scales = [1.]
    L = _libdc.lib()
    stream = C.c_void_p(_caffe._caffe.lib.caffe_stream())
    h, w = image.shape[:2]
    best_pose, highest_confidence = None, 0.
    out = np.zeros((5, 14), np.float32)
    d_pose = _POSE.reserve(out.nbytes)
    for scale_factor in scales:
        _, out_h, out_w, _ = _plan(h, w, scale_factor)
        net = _model(model_def, model_bin, out_h, out_w, weights)
        preprocess_to_blob(image, scale_factor, net.blobs['data'], stream)
        net.forward()
        prob, loc = net.blobs['prob'], net.blobs['loc_pred']
        _libdc.check(L.dc_pose_from_maps(prob.gpu_data_ptr(), loc.gpu_data_ptr(), 1, 14, prob.shape[2], prob.shape[3], _STRIDE,
                                         _LOCREF_SCALE_MUL, float(scale_factor), d_pose, stream))
        _libdc.check(L.dc_memcpy_async(out.ctypes.data_as(C.c_void_p), d_pose, out.nbytes, 2, stream))
        _libdc.check(L.dc_stream_sync(stream))
        pose = out.copy()
        minconf = float(pose[2].min())
        if minconf > highest_confidence:               # :121-126
            highest_confidence, best_pose = minconf, pose
    return best_pose
