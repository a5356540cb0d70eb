#!/usr/bin/env python
"""Kernel micro-benchmark for dc_conv_forward (tuning aid, not a bench number): times one conv
geometry with CUDA events, optionally with the residual epilogue."""
import argparse
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
libdc = importlib.import_module("deepcut-cnn_b200.libdc")
L = libdc.lib()
libdc.check(L.dc_init(0))


def run(n, h, w, cin, cout, k, pad, dil, residual, mode, iters=20):
    rows = L.dc_packed_rows(cout)
    K = k * k * cin
    x = torch.randn(2, n, h, w, cin, device="cuda").half()
    wp = torch.randn(2, rows, K, device="cuda").half()
    sc = torch.ones(rows, device="cuda")
    sh = torch.zeros(rows, device="cuda")
    ho, wo = h + 2 * pad - (dil * (k - 1) + 1) + 1, w + 2 * pad - (dil * (k - 1) + 1) + 1
    res = torch.randn(2, n, ho, wo, cout, device="cuda").half() if residual else None
    if mode == 0:
        out = torch.empty(2, n, ho, wo, cout, device="cuda", dtype=torch.half)
        ldc = 0
    elif mode == 1:
        out = torch.empty(n * ho * wo, rows, device="cuda")
        ldc = rows
    else:
        ldc = (n * ho * wo + 31) // 32 * 32
        out = torch.empty(rows, ldc, device="cuda")
    a = libdc.ConvArgs(x=x.data_ptr(), n=n, h=h, w=w, cin=cin, cout=cout, kh=k, kw=k, pad=pad, dilation=dil,
                       w_packed=wp.data_ptr(), scale=sc.data_ptr(), shift=sh.data_ptr(),
                       residual=res.data_ptr() if residual else None, relu=1 if mode == 0 else 0, out_f32_rows=mode, ldc=ldc,
                       out=out.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        libdc.check(L.dc_conv_forward(C.byref(a), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        libdc.check(L.dc_conv_forward(C.byref(a), st))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * n * ho * wo * cout * K
    by = 4.0 * (n * h * w * cin + n * ho * wo * cout * (2 if residual else 1))
    print("n%d %dx%d %d->%d k%d d%d res=%d mode=%d : %.1f us  %.0f TF/s (x3 issued %.0f)  %.0f GB/s" %
          (n, h, w, cin, cout, k, dil, residual, mode, ms * 1e3, fl / ms / 1e9, 3 * fl / ms / 1e9, by / ms / 1e6))


def run_latency(n, h, w, cin, cout, k, pad, dil, residual, chain=50, reps=10):
    """Latency regime: `chain` dependent launches of one small conv captured into a CUDA graph (no host launch cost, programmatic
    dependent launch between them, every launch waits for its predecessor like consecutive layers do); us per launch."""
    rows = L.dc_packed_rows(cout)
    K = k * k * cin
    x = torch.randn(2, n, h, w, cin, device="cuda").half()
    wp = torch.randn(2, rows, K, device="cuda").half()
    sc = torch.ones(rows, device="cuda")
    sh = torch.zeros(rows, device="cuda")
    ho, wo = h + 2 * pad - (dil * (k - 1) + 1) + 1, w + 2 * pad - (dil * (k - 1) + 1) + 1
    res = torch.randn(2, n, ho, wo, cout, device="cuda").half() if residual else None
    out = torch.empty(2, n, ho, wo, cout, device="cuda", dtype=torch.half)
    ws = torch.empty(L.dc_splitk_workspace_bytes(), dtype=torch.uint8, device="cuda")
    a = libdc.ConvArgs(x=x.data_ptr(), n=n, h=h, w=w, cin=cin, cout=cout, kh=k, kw=k, pad=pad, dilation=dil,
                       w_packed=wp.data_ptr(), scale=sc.data_ptr(), shift=sh.data_ptr(),
                       residual=res.data_ptr() if residual else None, relu=1, out_f32_rows=0, ldc=0, out=out.data_ptr(),
                       splitk_workspace=ws.data_ptr(), splitk_workspace_bytes=ws.numel())
    st = C.c_void_p()
    libdc.check(L.dc_stream_create(C.byref(st)))
    libdc.check(L.dc_conv_forward(C.byref(a), st))
    libdc.check(L.dc_stream_sync(st))
    g = C.c_void_p()
    libdc.check(L.dc_graph_begin(st))
    for _ in range(chain):
        libdc.check(L.dc_conv_forward(C.byref(a), st))
    libdc.check(L.dc_graph_end(st, C.byref(g)))
    e0, e1 = C.c_void_p(), C.c_void_p()
    libdc.check(L.dc_event_create(C.byref(e0)))
    libdc.check(L.dc_event_create(C.byref(e1)))
    libdc.check(L.dc_graph_launch(g, st))
    libdc.check(L.dc_stream_sync(st))
    libdc.check(L.dc_event_record(e0, st))
    for _ in range(reps):
        libdc.check(L.dc_graph_launch(g, st))
    libdc.check(L.dc_event_record(e1, st))
    ms = C.c_float()
    libdc.check(L.dc_event_elapsed_ms(e0, e1, C.byref(ms)))
    L.dc_graph_destroy(g)
    print("latency n%d %dx%d %d->%d k%d d%d res=%d split_k<=%d from %d K-steps : %.2f us / launch" %
          (n, h, w, cin, cout, k, dil, residual, L.dc_get_split_k(), L.dc_get_split_k_min_steps(), ms.value * 1e3 / (chain * reps)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="res4")
    args = ap.parse_args()
    if args.set == "res4":
        for res in (1, 0):
            run(16, 45, 80, 256, 1024, 1, 0, 1, res, 0)
        run(16, 45, 80, 256, 1024, 1, 0, 1, 0, 1)
        run(16, 45, 80, 256, 1024, 1, 0, 1, 0, 2)
        run(16, 45, 80, 1024, 256, 1, 0, 1, 0, 0)
        run(16, 45, 80, 256, 256, 3, 1, 1, 0, 0)
        run(16, 45, 80, 512, 512, 3, 2, 2, 0, 0)
        run(16, 90, 160, 128, 512, 1, 0, 1, 1, 0)
        run(16, 180, 320, 64, 256, 1, 0, 1, 1, 0)
    elif args.set == "pairs":      # the CTA-pair (cta_group::2) launches of res3 / res4 / res5: 1x1 reduce, 1x1 expand + shortcut
        run(16, 90, 160, 512, 128, 1, 0, 1, 0, 0)
        run(16, 90, 160, 128, 512, 1, 0, 1, 1, 0)
        run(16, 45, 80, 1024, 256, 1, 0, 1, 0, 0)
        run(16, 45, 80, 256, 1024, 1, 0, 1, 1, 0)
        run(16, 45, 80, 256, 1024, 1, 0, 1, 0, 0)
        run(16, 45, 80, 2048, 512, 1, 0, 1, 0, 0)
        run(16, 45, 80, 512, 2048, 1, 0, 1, 1, 0)
        run(16, 45, 80, 512, 2048, 1, 0, 1, 0, 0)
        run(16, 45, 80, 256, 256, 3, 1, 1, 0, 0)
        run(16, 45, 80, 512, 512, 3, 2, 2, 0, 0)
    elif args.set == "res2":
        run(16, 180, 320, 64, 64, 3, 1, 1, 0, 0)
        run(16, 180, 320, 64, 64, 1, 0, 1, 0, 0)
        run(16, 180, 320, 256, 64, 1, 0, 1, 0, 0)
        run(16, 180, 320, 64, 256, 1, 0, 1, 0, 0)
        run(16, 90, 160, 128, 128, 3, 1, 1, 0, 0)
        run(16, 90, 160, 512, 128, 1, 0, 1, 0, 0)
    elif args.set == "lat":        # one 512x512 image: res4 2b / 2a / 2c, res5 2b / 2a, res3 2b / 2a
        run_latency(1, 32, 32, 256, 256, 3, 1, 1, 0)
        run_latency(1, 32, 32, 1024, 256, 1, 0, 1, 0)
        run_latency(1, 32, 32, 256, 1024, 1, 0, 1, 1)
        run_latency(1, 32, 32, 512, 512, 3, 2, 2, 0)
        run_latency(1, 32, 32, 2048, 512, 1, 0, 1, 0)
        run_latency(1, 64, 64, 128, 128, 3, 1, 1, 0)
        run_latency(1, 64, 64, 512, 128, 1, 0, 1, 0)
    elif args.set == "bound":
        # what bounds the 1x1 convs: the same launches with the weight tiles (DC_DEBUG_SKIP=1), the activation tiles (2) or both (3)
        # no longer loaded after the first pipeline fill (results are garbage; only the time counts), and without the residual
        for skip in ("0", "1", "2", "3"):
            os.environ["DC_DEBUG_SKIP"] = skip
            print("--- DC_DEBUG_SKIP=%s (1: no weight-tile loads, 2: no activation-tile loads)" % skip)
            run(16, 45, 80, 256, 1024, 1, 0, 1, 1, 0)        # res4 2c + shortcut
            run(16, 45, 80, 256, 1024, 1, 0, 1, 0, 0)        # ... without the residual stream
            run(16, 45, 80, 1024, 256, 1, 0, 1, 0, 0)        # res4 2a
            run(16, 45, 80, 256, 256, 3, 1, 1, 0, 0)         # res4 2b
            run(16, 90, 160, 128, 512, 1, 0, 1, 1, 0)        # res3 2c
            run(16, 180, 320, 64, 256, 1, 0, 1, 1, 0)        # res2 2c
        os.environ.pop("DC_DEBUG_SKIP")
    elif args.set == "c3":
        run(16, 180, 320, 64, 64, 3, 1, 1, 0, 0)
        run(16, 90, 160, 128, 128, 3, 1, 1, 0, 0)
        run(16, 45, 80, 256, 256, 3, 1, 1, 0, 0)
        run(16, 45, 80, 512, 512, 3, 2, 2, 0, 0)
