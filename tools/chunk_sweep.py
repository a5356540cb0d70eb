#!/usr/bin/env python
"""Sweeps the fused plan's schedule knobs on one GPU in ONE process (weights packed once): for every configuration
(environment variables the planner reads when it (re)builds: DC_CHUNK_PLAN, DC_L2_CHUNK_MB, DC_INPLACE_RESIDUAL, ...)
times K forwards with CUDA events after warm-up, takes the per-step table, and compares the three outputs bitwise with the
first configuration's.  One JSON line per configuration on stdout / into --out.

  python tools/chunk_sweep.py --batch 16 --height 720 --width 1280 --out gpurun_out/sweep.jsonl \
      --config base:DC_CHUNK_PLAN=0,0,0,0 --config c243:DC_CHUNK_PLAN=0,2,4,2 ...
"""
import argparse
import collections
import ctypes as C
import importlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deepcut-cnn_b200", "python"))


def stage_of(name):
    m = re.match(r"res(\d)[a-z]\d*_branch(\w+)", name)
    if not m:
        return name.split("/")[0]
    b = m.group(2)
    return "res%s_%s" % (m.group(1), "b1" if b.startswith("1") else b[:2])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--model", default="152")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", action="append", default=[], help="name:VAR=value;VAR=value  (';' separates variables)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--describe", action="store_true", help="print each configuration's plan summary to stderr")
    args = ap.parse_args()
    import numpy as np
    import caffe
    libdc = importlib.import_module("deepcut-cnn_b200.libdc")
    synth = importlib.import_module("deepcut-cnn_b200.synth")
    ptx = importlib.import_module("deepcut-cnn_b200.prototxt")
    gen = importlib.import_module("deepcut-cnn_b200.gen_prototxt")
    L = libdc.lib()
    caffe.set_mode_gpu()
    caffe.set_device(0)
    stream = C.c_void_p(caffe._caffe.lib.caffe_stream())
    d = os.path.join(ROOT, "models", "_gen")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "sweep_resnet%s_%dx%d.prototxt" % (args.model, args.height, args.width))
    gen.write(path, stages=gen.STAGES_152 if args.model == "152" else gen.STAGES_101, height=args.height, width=args.width)
    net = caffe.Net(path, caffe.TEST)
    net.set_params(synth.calibrated_weights(ptx.parse_file(path)))
    net.blobs["data"].reshape(args.batch, 3, args.height, args.width)
    net.blobs["data"].data[...] = synth.images(args.batch, args.height, args.width)
    touched = set()
    first = None
    out = open(args.out, "a") if args.out else None
    for spec in args.config or ["default:"]:
        name, _, rest = spec.partition(":")
        for v in touched:
            os.environ.pop(v, None)
        for kv in filter(None, rest.split(";")):
            k, _, v = kv.partition("=")
            os.environ[k] = v
            touched.add(k)
        net.materialize_intermediates(True)      # drops the plan ...
        net.materialize_intermediates(False)     # ... so the next forward re-plans under this environment
        if args.describe:
            sys.stderr.write("[%s]\n%s\n" % (name, "\n".join(l for l in net.describe_plan().split("\n") if not l.startswith("  ConvBN"))))
        res = net.forward()
        assert net.fused_last_forward, net.fusion_diagnostic
        got = {k: np.array(v) for k, v in res.items()}
        if first is None:
            first = got
        diff = {k: float(np.abs(got[k].astype(np.float64) - first[k]).max()) for k in got}
        for _ in range(args.warmup - 1):
            net.forward()
        e0, e1 = C.c_void_p(), C.c_void_p()
        libdc.check(L.dc_event_create(C.byref(e0)))
        libdc.check(L.dc_event_create(C.byref(e1)))
        caffe.sync()
        l0 = L.dc_launch_count()
        libdc.check(L.dc_event_record(e0, stream))
        for _ in range(args.steps):
            net.forward()
        libdc.check(L.dc_event_record(e1, stream))
        caffe.sync()
        ms = C.c_float()
        libdc.check(L.dc_event_elapsed_ms(e0, e1, C.byref(ms)))
        launches = (L.dc_launch_count() - l0) // args.steps
        net.set_step_timing(True)
        for _ in range(3):
            net.forward()
        caffe.sync()
        per = collections.OrderedDict()
        for typ, sname, sms, fl, by in net.step_info():
            per[stage_of(sname)] = per.get(stage_of(sname), 0.0) + sms
        net.set_step_timing(False)
        line = {"config": name, "env": rest, "ms_per_step": ms.value / args.steps, "images_per_s": args.batch * args.steps / (ms.value / 1e3),
                "launches": int(launches), "arena_mib": net.arena_bytes >> 20, "max_abs_diff_vs_first": diff,
                "stage_ms": {k: round(v, 4) for k, v in per.items()}, "stage_total_ms": round(sum(per.values()), 4)}
        s = json.dumps(line)
        print(s, flush=True)
        if out:
            out.write(s + "\n")
            out.flush()


if __name__ == "__main__":
    main()
