#!/usr/bin/env python
"""Single-person pose demo on the B200 path: the pipeline of the reference's python/pose/pose_demo.py +
estimate_pose.py restated against the pycaffe-compatible shim (the reference files themselves are Python-2
/ scipy<1.2 code and are not on the GPU box).  Pre-processing follows estimate_pose.py:83-106 (edge-replicate
pad 64, bilinear rescale, mean subtraction, crop to a multiple of the stride) and runs on the device
(dc_preprocess_u8_forward, bit-exact with Pillow's resize); the 720p-and-larger images the reference tiles into
<= 700 px pieces (:160-221, a 2016 GPU-memory workaround) run whole; the read-out (:131-143) runs on the device
(dc_pose_from_maps).  All of it lives in deepcut-cnn_b200/python/pose/estimate_pose.py, which keeps the reference's
estimate_pose(image, model_def, model_bin, scales) signature.

  python tools/pose_demo_b200.py --model models/_gen/ResNet-152.prototxt [--weights X.caffemodel] [--image img.png] [--scales 1.0,0.8]
"""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deepcut-cnn_b200", "python"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", required=True)
    ap.add_argument("--weights", default=None)
    ap.add_argument("--image", default=None, help="image file (PIL); default: a seeded synthetic 720p image")
    ap.add_argument("--scales", default="1.0")
    ap.add_argument("--gpu", type=int, default=0)
    args = ap.parse_args()
    import caffe
    from pose.estimate_pose import estimate_pose
    synth = importlib.import_module("deepcut-cnn_b200.synth")
    ptx = importlib.import_module("deepcut-cnn_b200.prototxt")
    caffe.set_mode_gpu()
    caffe.set_device(args.gpu)
    weights = None
    if not args.weights:                  # no trained weights ship with the reference
        weights = synth.calibrated_weights(ptx.parse_file(args.model))
    if args.image:
        from PIL import Image
        img = np.asarray(Image.open(args.image).convert("RGB"))[:, :, ::-1]      # RGB -> BGR, pose_demo.py:116-121
    else:
        img = np.random.default_rng(20160505).integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    pose = estimate_pose(img, args.model, args.weights, [float(s) for s in args.scales.split(",")], weights=weights)
    np.set_printoptions(precision=2, suppress=True)
    print("pose (rows: x, y, confidence, offset_y, offset_x; 14 joints):")
    print(pose)


if __name__ == "__main__":
    main()
