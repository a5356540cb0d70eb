"""Independent PyTorch model of a deploy prototxt (TEST INFRASTRUCTURE).

Built on torch.nn.functional ops instead of im2col+GEMM so a mis-reading shared
with oracle/caffe_ref.py cannot hide ("unpinned" rows).
``quant`` optionally emulates the storage precision of the CUDA path (activations
and weights rounded to fp16, or to an fp16 hi+lo pair) to budget error on CPU.
"""
import torch
import torch.nn.functional as F

from oracle import prototxt as pt


def _rep(p, n, d):
    v = p.get(n)
    return v[0] if v else d


def q_fp16(t):
    return t.half().to(t.dtype)


def q_fp16x2(t):
    hi = t.half().to(t.dtype)
    lo = (t - hi).half().to(t.dtype)
    return hi + lo


def q_bf16x2(t):
    hi = t.bfloat16().to(t.dtype)
    lo = (t - hi).bfloat16().to(t.dtype)
    return hi + lo


def forward(net_param, params, x, dtype=torch.float64, quant=None, fold=False, want=()):
    """params: {layer: [np arrays]}; x: np NCHW.  quant: fn applied to every stored
    activation and to conv weights (None = exact).  Returns dict of output blobs."""
    blobs = {"data": torch.from_numpy(x).to(dtype)}
    layers = net_param["layer"]
    consumed = set()
    keep = {}
    P = {k: [torch.from_numpy(a).to(dtype) for a in v] for k, v in params.items()}
    qa = quant or (lambda t: t)
    if quant:
        blobs["data"] = blobs["data"]
    for l in layers:
        t, name = pt.get(l, "type"), pt.get(l, "name")
        bots = [blobs[b] for b in l.get("bottom", [])]
        consumed.update(l.get("bottom", []))
        top = l["top"][0]
        if t == "Convolution":
            cp = pt.get(l, "convolution_param")
            w = P[name][0]
            b = P[name][1] if len(P[name]) > 1 else None
            y = F.conv2d(qa(bots[0]), qa(w), b, _rep(cp, "stride", 1), _rep(cp, "pad", 0), _rep(cp, "dilation", 1))
        elif t == "Deconvolution":
            cp = pt.get(l, "convolution_param")
            w = P[name][0]
            b = P[name][1] if len(P[name]) > 1 else None
            y = F.conv_transpose2d(qa(bots[0]), qa(w), b, _rep(cp, "stride", 1), _rep(cp, "pad", 0))
        elif t == "BatchNorm":
            m, v, sf = P[name]
            sf = 0.0 if float(sf[0]) == 0 else 1.0 / sf[0]
            eps = pt.get(pt.get(l, "batch_norm_param", {}), "eps", 1e-5)
            y = (bots[0] - (m * sf).view(1, -1, 1, 1)) / torch.sqrt(v * sf + eps).view(1, -1, 1, 1)
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
            y = F.max_pool2d(bots[0], pt.get(pp, "kernel_size"), pt.get(pp, "stride", 1),
                             pt.get(pp, "pad", 0), ceil_mode=True)
        elif t == "Crop":
            y = bots[0][:, :, :bots[1].shape[2], :bots[1].shape[3]]
        elif t == "Sigmoid":
            y = torch.sigmoid(bots[0])
        else:
            raise NotImplementedError(t)
        blobs[top] = y
        if name in want:
            keep[name] = y
    outs = {k: v for k, v in blobs.items() if k not in consumed}
    outs.update(keep)
    return {k: v.to(torch.float64).numpy() for k, v in outs.items()}
