#!/usr/bin/env python
"""`caffe time` for the B200 host (reference tools/caffe.cpp:302-388: time every layer's Forward over N
iterations).  Here the unit is the fused step; per-step times come from CUDA events on the forward stream.

  python tools/caffe_time.py --model models/_gen/ResNet-152.prototxt --batch 16 --height 720 --width 1280 --iterations 10
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deepcut-cnn_b200", "python"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", required=True, help="deploy prototxt")
    ap.add_argument("--weights", default=None, help=".caffemodel (default: constant fillers)")
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--per-layer", action="store_true", help="time the per-layer plugin path instead of the fused plan")
    args = ap.parse_args()
    import time
    import numpy as np
    import caffe
    caffe.set_mode_gpu()
    caffe.set_device(args.gpu)
    net = caffe.Net(args.model, args.weights, caffe.TEST) if args.weights else caffe.Net(args.model, caffe.TEST)
    data = net.blobs[net.inputs[0]]
    n, c, h, w = data.shape
    data.reshape(args.batch or n, c, args.height or h, args.width or w)
    data.data[...] = np.random.default_rng(0).standard_normal(data.shape).astype(np.float32)
    if args.per_layer:
        net.set_fusion(False)
    net.forward()
    caffe.sync()
    print("*** Benchmark begins ***  Testing for %d iterations.  fused=%s  launches/forward=%d" %
          (args.iterations, net.fused_last_forward, net.last_forward_launches))
    if not net.fused_last_forward:
        t0 = time.time()
        for _ in range(args.iterations):
            net.forward()
        caffe.sync()
        print("Average Forward pass: %.3f ms (per-layer path; %s)" % ((time.time() - t0) * 1e3 / args.iterations, net.fusion_diagnostic or "fusion off"))
        return
    net.set_step_timing(True)
    acc = collections.OrderedDict()
    for _ in range(args.iterations):
        net.forward()
        caffe.sync()
        for typ, name, ms, fl, by in net.step_info():
            a = acc.setdefault(name, [typ, 0.0, fl, by])
            a[1] += ms
    total = 0.0
    for name, (typ, ms, fl, by) in acc.items():
        ms /= args.iterations
        total += ms
        print("%-12s %-40s forward: %8.4f ms  %7.1f TFLOP/s  %7.1f GB/s" % (typ, name[:40], ms, fl / ms / 1e9 if ms else 0, by / ms / 1e6 if ms else 0))
    print("Average Forward pass: %.3f ms.  (%d fused steps, arena %d MiB, packed weights %d MiB)" %
          (total, len(acc), net.arena_bytes >> 20, net.weight_bytes >> 20))
    print("*** Benchmark ends ***")


if __name__ == "__main__":
    main()
