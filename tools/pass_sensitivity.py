#!/usr/bin/env python
"""Per-layer-group sensitivity of the outputs to the number of tensor-core products (CPU study, fp64 emulation).

The CUDA path stores every operand as an fp16 pair (hi = fp16(x), lo = fp16(x - hi)) and issues three products per K-step:
hi*hi + hi*lo + lo*hi.  Dropping `lo*hi` means the ACTIVATIONS of that layer are effectively single fp16; dropping `hi*lo`
means its WEIGHTS are; dropping both is the plain one-pass fp16 GEMM.  This tool answers VERDICT r1 item 3: is there a set of
layers that tolerates fewer products within the parity budget (1e-3 max-abs on prob / loc_pred / next_pred, target 5e-4)?

For each layer group (stage x {branch1, 2a, 2b, 2c}, stem, heads) and each reduced variant it evaluates the shipped
ResNet-152 deploy net (calibrated synthetic weights, the parity tests' recipe) in fp64 with exactly that group reduced and
every other layer at the full three products, and reports the max-abs deviation of the three outputs from the exact fp64
forward.  Accumulation is exact (fp64), so the numbers isolate OPERAND precision; the GPU adds its fp32 accumulation on top.

  python tools/pass_sensitivity.py --size 256 --out profiles/r2_pass_sensitivity.json
"""
import argparse
import importlib
import json
import os
import re
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.nn.functional as F

from oracle import prototxt as pt


def split(t):
    hi = t.half().to(t.dtype)
    lo = (t - hi).half().to(t.dtype)
    return hi, lo


def group_of(name):
    m = re.match(r"res(\d)[a-z]\d*_branch(\w+)", name)
    if m:
        b = m.group(2)
        return "res%s_%s" % (m.group(1), "b1" if b.startswith("1") else b[:2])
    if name == "conv1":
        return "conv1"
    return "heads"


def forward(net_param, P, x, variant_of):
    """variant_of(layer name) -> 'full' (3 products) | 'act16' (no lo*hi) | 'w16' (no hi*lo) | 'one' (hi*hi only) | 'exact'."""
    blobs = {"data": x}
    consumed = set()
    for l in net_param["layer"]:
        t, name = pt.get(l, "type"), pt.get(l, "name")
        bots = [blobs[b] for b in l.get("bottom", [])]
        consumed.update(l.get("bottom", []))
        top = l["top"][0]
        if t in ("Convolution", "Deconvolution"):
            cp = pt.get(l, "convolution_param")
            w = P[name][0]
            b = P[name][1] if len(P[name]) > 1 else None
            rep = lambda n, d: (cp.get(n) or [d])[0]
            if t == "Convolution":
                op = lambda a, ww, bias=None: F.conv2d(a, ww, bias, rep("stride", 1), rep("pad", 0), rep("dilation", 1))
            else:
                op = lambda a, ww, bias=None: F.conv_transpose2d(a, ww, bias, rep("stride", 1), rep("pad", 0))
            v = variant_of(name)
            if v == "exact":
                y = op(bots[0], w, b)
            else:
                xh, xl = split(bots[0])
                # the weights' per-row power-of-two scaling (dc_pack_conv_weight) does not change an fp16 split's relative precision
                wh, wl = split(w)
                y = op(xh, wh, b)
                if v in ("full", "act16"):
                    y = y + op(xh, wl)
                if v in ("full", "w16"):
                    y = y + op(xl, wh)
        elif t == "BatchNorm":
            m, var, sf = P[name]
            sf = 0.0 if float(sf[0]) == 0 else 1.0 / sf[0]
            y = (bots[0] - (m * sf).view(1, -1, 1, 1)) / torch.sqrt(var * sf + 1e-5).view(1, -1, 1, 1)
        elif t == "Scale":
            y = bots[0] * P[name][0].view(1, -1, 1, 1)
            if len(P[name]) > 1:
                y = y + P[name][1].view(1, -1, 1, 1)
        elif t == "ReLU":
            y = torch.relu(bots[0])
        elif t == "Eltwise":
            y = bots[0] + bots[1]
        elif t == "Pooling":
            pp = pt.get(l, "pooling_param")
            y = F.max_pool2d(bots[0], pt.get(pp, "kernel_size"), pt.get(pp, "stride", 1), pt.get(pp, "pad", 0), ceil_mode=True)
        elif t == "Crop":
            y = bots[0][:, :, :bots[1].shape[2], :bots[1].shape[3]]
        elif t == "Sigmoid":
            y = torch.sigmoid(bots[0])
        else:
            raise NotImplementedError(t)
        blobs[top] = y
    return {k: v for k, v in blobs.items() if k not in consumed}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_pass_sensitivity.json"))
    ap.add_argument("--variants", default="act16,w16,one")
    args = ap.parse_args()
    gen = importlib.import_module("deepcut-cnn_b200.gen_prototxt")
    synth = importlib.import_module("deepcut-cnn_b200.synth")
    ptx = importlib.import_module("deepcut-cnn_b200.prototxt")
    path = "/tmp/pass_sens_%d.prototxt" % args.size
    gen.write(path, height=args.size, width=args.size)
    weights = synth.calibrated_weights(ptx.parse_file(path))
    net_param = pt.parse_file(path)
    P = {k: [torch.from_numpy(np.asarray(a)).double() for a in v] for k, v in weights.items()}
    x = torch.from_numpy(synth.images(1, args.size, args.size)).double()
    convs = [pt.get(l, "name") for l in net_param["layer"] if pt.get(l, "type") in ("Convolution", "Deconvolution")]
    groups = []
    flops = {}
    for n in convs:
        g = group_of(n)
        if g not in groups:
            groups.append(g)
    outs = ("prob", "loc_pred", "next_pred")
    t0 = time.time()
    exact = forward(net_param, P, x, lambda n: "exact")
    full = forward(net_param, P, x, lambda n: "full")
    err = lambda a: {k: float((a[k] - exact[k]).abs().max()) for k in outs}
    doc = {"net": "ResNet-152 deploy, calibrated synthetic weights", "input": "1x3x%dx%d" % (args.size, args.size),
           "arith": "fp64 accumulation; operands split into fp16 hi+lo as on the GPU", "budget": 1e-3,
           "all_layers_full_3_products": err(full), "layers_per_group": {g: sum(1 for n in convs if group_of(n) == g) for g in groups},
           "groups": {}}
    print("exact + full: %.1f s; full-3-product error %s" % (time.time() - t0, doc["all_layers_full_3_products"]), flush=True)
    for v in args.variants.split(","):
        allv = forward(net_param, P, x, lambda n, v=v: v)
        doc.setdefault("all_layers", {})[v] = err(allv)
        print("ALL layers %-6s %s" % (v, doc["all_layers"][v]), flush=True)
    for g in groups:
        doc["groups"][g] = {}
        for v in args.variants.split(","):
            got = forward(net_param, P, x, lambda n, g=g, v=v: v if group_of(n) == g else "full")
            doc["groups"][g][v] = err(got)
            print("%-10s %-6s %s" % (g, v, "  ".join("%s %.2e" % kv for kv in doc["groups"][g][v].items())), flush=True)
        json.dump(doc, open(args.out, "w"), indent=1)
    json.dump(doc, open(args.out, "w"), indent=1)
    print("done in %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
