"""Seeded synthetic inputs and weights for the DeeperCut deploy net.

The reference ships no trained weights (models/deepercut/download_models.sh needs
the network) and the deploy prototxt names no fillers, so an unloaded net is all
zeros (caffe.proto:43-46).  This is the recipe of SURVEY.md section 8(d): it keeps
activations O(1) through the 50 residual adds.  Used by tests, bench.py and
smoke(); it is host-side harness code, not part of the hot path.
"""
import numpy as np

MEAN_BGR = (104.0, 117.0, 123.0)   # python/pose/estimate_pose.py:25


def images_u8(n, h, w, seed=20160505):
    """The decoded images themselves: uint8 [n][h][w][3] (what images() subtracts the mean from)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)


def images(n, h, w, seed=20160505):
    """uint8 image -> float32 NCHW minus the demo's per-channel mean."""
    u8 = images_u8(n, h, w, seed)
    x = u8.astype(np.float32) - np.asarray(MEAN_BGR, np.float32)
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))


def weights(param_shapes, seed=152):
    """param_shapes: ordered {layer_name: (layer_type, [blob shapes])} in prototxt
    order.  Returns {layer_name: [float32 arrays]} in the reference's blob orders
    (conv: W[,b]; BatchNorm: mean_sum, var_sum, scale_factor; Scale: gamma, beta)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    f32 = np.float32
    for name, (ltype, shapes) in param_shapes.items():
        blobs = []
        if ltype in ("Convolution", "Deconvolution"):
            ws = shapes[0]
            is_head = len(shapes) > 1          # only the 6 head layers carry a bias
            if is_head:
                blobs.append(rng.normal(0.0, 0.01, ws).astype(f32))
                blobs.append(rng.normal(0.0, 0.1, shapes[1]).astype(f32))
            else:
                fan_in = ws[1] * ws[2] * ws[3]   # MSRAFiller, filler.hpp:186-241
                blobs.append(rng.normal(0.0, np.sqrt(2.0 / fan_in), ws).astype(f32))
        elif ltype == "BatchNorm":
            c = shapes[0]
            blobs.append(rng.normal(0.0, 0.1, c).astype(f32))
            blobs.append(rng.uniform(0.5, 1.5, c).astype(f32))
            blobs.append(np.ones(1, f32))
        elif ltype == "Scale":
            c = shapes[0]
            lo, hi = (0.1, 0.3) if name.endswith("_branch2c") else (0.8, 1.2)
            blobs.append(rng.uniform(lo, hi, c).astype(f32))
            if len(shapes) > 1:
                blobs.append(rng.normal(0.0, 0.05, shapes[1]).astype(f32))
        else:
            continue
        out[name] = blobs
    return out


# --------------------------------------------------------------------------
# Trained-like weights: BatchNorm statistics calibrated to the activations
# --------------------------------------------------------------------------
def _rep(msg, name, default):
    v = msg.get(name)
    return v[0] if v else default


def _coarse(t):
    """Keep 8 mantissa bits so the result is identical on any host/thread count."""
    import torch
    return t.to(torch.bfloat16).to(torch.float64)


def calibrated_weights(net, seed=152, calib_hw=(128, 128)):
    """Like ``weights`` but every BatchNorm's stored statistics are set from the
    actual per-channel mean/variance of its input on a seeded calibration image
    (then perturbed), the way a trained net's running statistics relate to its
    activations.  Without this the random net's activations grow geometrically
    through the 50 residual blocks (|res5c| ~ 1e3-1e4, logits ~ 1e3, prob
    saturated) and an absolute 1e-3 tolerance on the outputs is meaningless.

    net: parsed NetParameter (``prototxt.parse``).  The calibration pass runs in
    fp64 on the CPU with torch -- harness code, executed once per weight set.
    Returns {layer_name: [float32 arrays]} in the reference's blob orders.
    """
    import torch
    import torch.nn.functional as F
    rng = np.random.Generator(np.random.PCG64(seed))
    f32 = np.float32
    x = torch.from_numpy(images(1, calib_hw[0], calib_hw[1], seed=seed + 1)).double()
    blobs = {net.one("input", "data"): x}
    out = {}
    for l in net.rep("layer"):
        t, name = l.one("type"), l.one("name")
        bots = [blobs[b] for b in l.rep("bottom")]
        top = l.rep("top")[0]
        if t in ("Convolution", "Deconvolution"):
            cp = l.sub("convolution_param")
            co, k = cp.one("num_output"), _rep(cp, "kernel_size", 1)
            s, p, d = _rep(cp, "stride", 1), _rep(cp, "pad", 0), _rep(cp, "dilation", 1)
            ci = bots[0].shape[1]
            has_bias = cp.one("bias_term", True)
            ws = (co, ci, k, k) if t == "Convolution" else (ci, co, k, k)
            if has_bias:     # the 6 head layers
                w = rng.normal(0.0, 0.01, ws).astype(f32)
                b = rng.normal(0.0, 0.1, (co,)).astype(f32)
                out[name] = [w, b]
            else:
                w = rng.normal(0.0, np.sqrt(2.0 / (ci * k * k)), ws).astype(f32)
                b = None
                out[name] = [w]
            wt = torch.from_numpy(w).double()
            bt = torch.from_numpy(b).double() if b is not None else None
            if t == "Convolution":
                y = F.conv2d(bots[0], wt, bt, s, p, d)
            else:
                y = F.conv_transpose2d(bots[0], wt, bt, s, p)
        elif t == "BatchNorm":
            c = bots[0].shape[1]
            m = bots[0].mean(dim=(0, 2, 3))
            v = bots[0].var(dim=(0, 2, 3), unbiased=False) + 1e-6
            mean = _coarse(m + v.sqrt() * torch.from_numpy(rng.normal(0.0, 0.1, c)))
            var = _coarse(v * torch.from_numpy(rng.uniform(0.5, 1.5, c)))
            out[name] = [mean.numpy().astype(f32), var.numpy().astype(f32), np.ones(1, f32)]
            eps = l.sub("batch_norm_param").one("eps", 1e-5)
            y = (bots[0] - mean.view(1, -1, 1, 1)) / torch.sqrt(var + eps).view(1, -1, 1, 1)
        elif t == "Scale":
            c = bots[0].shape[1]
            lo, hi = (0.1, 0.3) if name.endswith("_branch2c") else (0.8, 1.2)
            g = rng.uniform(lo, hi, c).astype(f32)
            out[name] = [g]
            y = bots[0] * torch.from_numpy(g).double().view(1, -1, 1, 1)
            if l.sub("scale_param").one("bias_term", False):
                b = rng.normal(0.0, 0.05, c).astype(f32)
                out[name].append(b)
                y = y + torch.from_numpy(b).double().view(1, -1, 1, 1)
        elif t == "ReLU":
            y = torch.relu(bots[0])
        elif t == "Eltwise":
            y = bots[0] + bots[1]
        elif t == "Pooling":
            pp = l.sub("pooling_param")
            y = F.max_pool2d(bots[0], pp.one("kernel_size"), pp.one("stride", 1), pp.one("pad", 0), ceil_mode=True)
        elif t == "Crop":
            y = bots[0][:, :, :bots[1].shape[2], :bots[1].shape[3]]
        elif t == "Sigmoid":
            y = torch.sigmoid(bots[0])
        else:
            raise NotImplementedError("synth: layer type %s" % t)
        blobs[top] = y
    return out
