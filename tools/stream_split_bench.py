#!/usr/bin/env python
"""Does running the batch as T concurrent sub-batch forwards (T Nets of batch B/T, one host thread + stream each, device-resident
inputs) fill the tail rounds of the persistent conv kernels?  Tuning experiment, not a bench number: prints ms per B images for
each T.  Every Net keeps its input on the device (no copies in the timed region)."""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deepcut-cnn_b200", "python"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--splits", default="1,2,4")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import caffe
    synth = importlib.import_module("deepcut-cnn_b200.synth")
    ptx = importlib.import_module("deepcut-cnn_b200.prototxt")
    gen = importlib.import_module("deepcut-cnn_b200.gen_prototxt")
    caffe.set_mode_gpu()
    caffe.set_device(0)
    path = os.path.join(ROOT, "models", "_gen", "split_bench_%dx%d.prototxt" % (args.height, args.width))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    gen.write(path, height=args.height, width=args.width)
    weights = synth.calibrated_weights(ptx.parse_file(path))
    x = synth.images(args.batch, args.height, args.width)
    for T in [int(v) for v in args.splits.split(",")]:
        if args.batch % T:
            continue
        b = args.batch // T
        ready, go, done = (threading.Barrier(T + 1) for _ in range(3))
        errors = []

        def worker(idx):
            try:
                caffe.set_mode_gpu()
                caffe.set_device(0)
                net = caffe.Net(path, caffe.TEST)
                net.set_params(weights)
                net.blobs["data"].reshape(b, 3, args.height, args.width)
                net.blobs["data"].data[...] = x[idx * b:(idx + 1) * b]
                for _ in range(3):
                    net.forward()
                caffe.sync()
                ready.wait()
                go.wait()
                for _ in range(args.steps):
                    net.forward()
                caffe.sync()
                done.wait()
                del net
            except Exception as exc:
                errors.append(exc)
                for bar in (ready, go, done):
                    bar.abort()

        threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(T)]
        for t in threads:
            t.start()
        ready.wait()
        t0 = time.time()
        go.wait()
        done.wait()
        ms = (time.time() - t0) * 1e3 / args.steps
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        rec = {"streams": T, "images_per_stream": b, "ms_per_%d_images" % args.batch: ms, "images_per_s": args.batch / ms * 1e3}
        print(json.dumps(rec), flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
