#!/usr/bin/env python
"""Single-person pose demo on the B200 path: the pipeline of the reference's python/pose/pose_demo.py +
estimate_pose.py restated against the pycaffe-compatible shim (the reference files themselves are Python-2
/ scipy<1.2 code and are not on the GPU box).  Pre-processing follows estimate_pose.py:83-106 (edge-replicate
pad 64, bilinear rescale, mean subtraction, crop to a multiple of the stride); the 720p-and-larger images the
reference tiles into <= 700 px pieces (:160-221, a 2016 GPU-memory workaround) run whole; the read-out
(:131-143) runs on the device (dc_pose_from_maps).

  python tools/pose_demo_b200.py --model models/_gen/ResNet-152.prototxt [--weights X.caffemodel] [--image img.png] [--scales 1.0,0.8]
"""
import argparse
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deepcut-cnn_b200", "python"))

MEAN = np.array([104., 117., 123.], np.float32)      # estimate_pose.py:25
STRIDE = 8.0
LOCREF_SCALE = float(np.sqrt(53.0))                   # estimate_pose.py:27


def bilinear_resize(img, factor):
    """scipy.misc.imresize(image, factor, interp='bilinear') stand-in (PIL BILINEAR on uint8, as imresize does)."""
    from PIL import Image
    h, w = img.shape[:2]
    size = (int(w * factor), int(h * factor))
    return np.asarray(Image.fromarray(img.astype(np.uint8)).resize(size, Image.BILINEAR))


def prepare(image_bgr, scale):
    """estimate_pose.py:83-106."""
    h, w = image_bgr.shape[:2]
    bg_w = int(np.ceil(float(w) * scale / STRIDE) * STRIDE)
    bg_h = int(np.ceil(float(h) * scale / STRIDE) * STRIDE)
    img = np.vstack((image_bgr, np.tile(image_bgr[-1:], (64, 1, 1))))
    img = np.hstack((img, np.tile(img[:, -1:], (1, 64, 1))))
    img = bilinear_resize(img, scale).astype(np.float32) - MEAN
    net_input = np.zeros((bg_h, bg_w, 3), np.float32)
    hh, ww = min(bg_h, img.shape[0]), min(bg_w, img.shape[1])
    net_input[:hh, :ww] = img[:hh, :ww]
    return net_input.transpose(2, 0, 1)


def estimate_pose(net, image_bgr, scales, libdc, stream):
    L = libdc.lib()
    best, best_conf = None, 0.0
    out = np.zeros((1, 5, 14), np.float32)
    for s in scales:
        x = prepare(image_bgr, s)
        net.blobs["data"].reshape(1, 3, x.shape[1], x.shape[2])
        net.blobs["data"].data[0, ...] = x
        net.forward()
        prob, loc = net.blobs["prob"], net.blobs["loc_pred"]
        d_out = C.c_void_p()
        libdc.check(L.dc_malloc(C.byref(d_out), out.nbytes))
        libdc.check(L.dc_pose_from_maps(prob.gpu_data_ptr(), loc.gpu_data_ptr(), 1, 14, prob.shape[2], prob.shape[3], STRIDE, LOCREF_SCALE,
                                        float(s), d_out, stream))
        libdc.check(L.dc_memcpy_async(out.ctypes.data_as(C.c_void_p), d_out, out.nbytes, 2, stream))
        libdc.check(L.dc_stream_sync(stream))
        L.dc_free(d_out)
        pose = out[0].copy()
        if pose[2].min() > best_conf:                 # estimate_pose.py:121-126: best minimum confidence wins
            best_conf, best = float(pose[2].min()), pose
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", required=True)
    ap.add_argument("--weights", default=None)
    ap.add_argument("--image", default=None, help="image file (PIL); default: a seeded synthetic 720p image")
    ap.add_argument("--scales", default="1.0")
    ap.add_argument("--gpu", type=int, default=0)
    args = ap.parse_args()
    import caffe
    libdc = importlib.import_module("deepcut-cnn_b200.libdc")
    synth = importlib.import_module("deepcut-cnn_b200.synth")
    ptx = importlib.import_module("deepcut-cnn_b200.prototxt")
    caffe.set_mode_gpu()
    caffe.set_device(args.gpu)
    if args.weights:
        net = caffe.Net(args.model, args.weights, caffe.TEST)
    else:
        net = caffe.Net(args.model, caffe.TEST)
        net.set_params(synth.calibrated_weights(ptx.parse_file(args.model)))      # no trained weights ship with the reference
    if args.image:
        from PIL import Image
        img = np.asarray(Image.open(args.image).convert("RGB"))[:, :, ::-1]      # RGB -> BGR, pose_demo.py:116-121
    else:
        img = np.random.default_rng(20160505).integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    stream = C.c_void_p(caffe._caffe.lib.caffe_stream())
    pose = estimate_pose(net, img, [float(s) for s in args.scales.split(",")], libdc, stream)
    np.set_printoptions(precision=2, suppress=True)
    print("pose (rows: x, y, confidence, offset_y, offset_x; 14 joints):")
    print(pose)


if __name__ == "__main__":
    main()
