#!/usr/bin/env python
"""Minimal forward loop for ncu (never a bench number): builds the bench workload, runs --warm
forwards, then --iters forwards.  One forward = 162 kernel launches at the default workload."""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deepcut-cnn_b200", "python"))

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--height", type=int, default=720)
ap.add_argument("--width", type=int, default=1280)
ap.add_argument("--warm", type=int, default=2)
ap.add_argument("--iters", type=int, default=1)
ap.add_argument("--list-steps", action="store_true")
ap.add_argument("--profiler-range", action="store_true", help="cudaProfilerStart/Stop around the --iters forwards (ncu --profile-from-start off)")
ap.add_argument("--schedule-out", default=None, help="write the plan description incl. the launch order (DC_DESCRIBE_SCHEDULE) here")
args = ap.parse_args()

import caffe  # noqa: E402
gen = importlib.import_module("deepcut-cnn_b200.gen_prototxt")
synth = importlib.import_module("deepcut-cnn_b200.synth")
ptx = importlib.import_module("deepcut-cnn_b200.prototxt")
caffe.set_mode_gpu()
caffe.set_device(0)
path = os.path.join(ROOT, "models", "_gen", "profile_%dx%d.prototxt" % (args.height, args.width))
os.makedirs(os.path.dirname(path), exist_ok=True)
gen.write(path, height=args.height, width=args.width)
net = caffe.Net(path, caffe.TEST)
net.set_params(synth.calibrated_weights(ptx.parse_file(path)))
net.blobs["data"].reshape(args.batch, 3, args.height, args.width)
net.blobs["data"].data[...] = synth.images(args.batch, args.height, args.width)
for _ in range(args.warm):
    net.forward()
caffe.sync()
if args.schedule_out:
    os.environ["DC_DESCRIBE_SCHEDULE"] = "1"
    open(args.schedule_out, "w").write(net.describe_plan())
if args.profiler_range:
    import torch
    torch.cuda.profiler.start()
for _ in range(args.iters):
    net.forward()
caffe.sync()
if args.profiler_range:
    torch.cuda.profiler.stop()
if args.list_steps:
    net.set_step_timing(True)
    net.forward()
    caffe.sync()
    for i, s in enumerate(net.step_info()):
        print(i, s[0], s[1], "%.3f ms" % s[2])
