"""Whole-net harness shared by the CPU and GPU tests: same prototxt + same seeded weights into the
oracle (oracle/caffe_ref.py) and into the product (the pycaffe-compatible shim over the C++ host)."""
import os

import numpy as np

import dcutil
from oracle import caffe_ref

_cache = {}


def build(tmpdir, stages=(1, 1, 1, 1), h=64, w=64):
    """-> (prototxt path, weights dict).  Weights are calibrated once per topology (fp64 CPU pass)."""
    key = (tuple(stages),)
    path = dcutil.write_prototxt(tmpdir, stages=tuple(stages), height=h, width=w)
    if key not in _cache:
        _cache[key] = dcutil.synth.calibrated_weights(dcutil.ptx.parse_file(path))
    return path, _cache[key]


def oracle_forward(path, weights, x, want=None):
    net = caffe_ref.load_net(path)
    net.reshape_input("data", x.shape)
    net.params = weights
    return net.forward({"data": x}, want=want)


def reference_available():
    """oracle/_ref: built here when /root/reference exists, prebuilt on the GPU box."""
    from oracle import build_ref, ref_caffe
    build_ref.build()
    return ref_caffe.available()


def reference_forward(path, weights, x, want=None):
    """The reference's own CPU layer code (oracle/_ref) on the same prototxt + weights."""
    from oracle import ref_caffe
    net = ref_caffe.RefCaffeNet(open(path).read())
    net.set_params(weights)
    return net.forward({"data": x}, want=want)


def product_net(path, weights, device=0):
    caffe = dcutil.caffe_module()
    caffe.set_mode_gpu()
    caffe.set_device(device)
    net = caffe.Net(path, caffe.TEST)
    net.set_params(weights)
    return net


def product_forward(net, x):
    net.blobs["data"].reshape(*x.shape)
    net.blobs["data"].data[...] = x
    out = net.forward()
    return {k: np.array(v) for k, v in out.items()}


def max_err(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())
