"""numpy restatement of the reference's CPU forward path (TEST INFRASTRUCTURE).

Every function cites the reference file:line (relative to /root/reference) whose
arithmetic it follows.  fp32 everywhere, NCHW, one image at a time through
im2col + SGEMM exactly like ``ConvolutionLayer::Forward_cpu``; the SGEMM is
numpy's (OpenBLAS ``cblas_sgemm``, one of the reference's three supported BLAS
back-ends, Makefile:361-363).  Nothing here is imported by the product.
"""
import math

import numpy as np

from . import prototxt as pt

F32 = np.float32


# --------------------------------------------------------------------------
# util/im2col.cpp, util/math_functions.cpp
# --------------------------------------------------------------------------
def conv_out_size(size, k, pad, stride, dil):
    """conv_layer.cpp:8-22  out = (in + 2p - (d(k-1)+1)) / s + 1."""
    return (size + 2 * pad - (dil * (k - 1) + 1)) // stride + 1


def im2col(x, kh, kw, ph, pw, sh, sw, dh, dw):
    """im2col_cpu, src/caffe/util/im2col.cpp:18-55.

    x: [C,H,W] -> col [C*kh*kw, Ho*Wo]; row index = c*kh*kw + p*kw + q,
    input_row = -pad + p*dilation + out_row*stride, zero outside the image.
    """
    C, H, W = x.shape
    Ho = conv_out_size(H, kh, ph, sh, dh)
    Wo = conv_out_size(W, kw, pw, sw, dw)
    col = np.zeros((C, kh, kw, Ho, Wo), F32)
    for p in range(kh):
        for q in range(kw):
            ry = -ph + p * dh + np.arange(Ho) * sh
            cx = -pw + q * dw + np.arange(Wo) * sw
            vy = np.nonzero((ry >= 0) & (ry < H))[0]
            vx = np.nonzero((cx >= 0) & (cx < W))[0]
            if len(vy) == 0 or len(vx) == 0:
                continue
            col[:, p, q, vy[0]:vy[-1] + 1, vx[0]:vx[-1] + 1] = x[
                :, ry[vy[0]]: ry[vy[-1]] + 1: sh, cx[vx[0]]: cx[vx[-1]] + 1: sw]
    return col.reshape(C * kh * kw, Ho * Wo), Ho, Wo


def col2im(col, C, H, W, kh, kw, ph, pw, sh, sw, dh, dw):
    """col2im_cpu, src/caffe/util/im2col.cpp:162-197 (zero, then += per (c,p,q)
    in that loop order, so each output element's fp32 adds happen in (p,q) order)."""
    Ho = conv_out_size(H, kh, ph, sh, dh)
    Wo = conv_out_size(W, kw, pw, sw, dw)
    col = col.reshape(C, kh, kw, Ho, Wo)
    im = np.zeros((C, H, W), F32)
    for p in range(kh):
        for q in range(kw):
            r0 = -ph + p * dh
            c0 = -pw + q * dw
            oy = np.arange(Ho)
            ox = np.arange(Wo)
            ry = r0 + oy * sh
            cx = c0 + ox * sw
            vy = (ry >= 0) & (ry < H)
            vx = (cx >= 0) & (cx < W)
            if not vy.any() or not vx.any():
                continue
            oy0, oy1 = np.nonzero(vy)[0][[0, -1]]
            ox0, ox1 = np.nonzero(vx)[0][[0, -1]]
            im[:, ry[oy0]: ry[oy1] + 1: sh, cx[ox0]: cx[ox1] + 1: sw] += \
                col[:, p, q, oy0:oy1 + 1, ox0:ox1 + 1]
    return im


def sgemm(a, b):
    """caffe_cpu_gemm -> cblas_sgemm(RowMajor), math_functions.cpp:12-21."""
    return np.matmul(np.ascontiguousarray(a, F32), np.ascontiguousarray(b, F32))


# --------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------
def _hw(v):
    return (v, v) if np.isscalar(v) else tuple(v)


def convolution(x, w, b, stride, pad, dil):
    """ConvolutionLayer::Forward_cpu conv_layer.cpp:25-40 ->
    forward_cpu_gemm base_conv_layer.cpp:256-272 (im2col skipped iff 1x1/s1/p0,
    :109-116) + forward_cpu_bias :274-280.  w: [Cout,Cin,kh,kw]; stride/pad/dil: int or (h, w)."""
    N, C, H, W = x.shape
    Co, Ci, kh, kw = w.shape
    assert Ci == C, "groups unsupported (deepercut uses group=1)"
    (sh, sw), (ph, pw), (dh, dw) = _hw(stride), _hw(pad), _hw(dil)
    is_1x1 = kh == 1 and kw == 1 and sh == 1 and sw == 1 and ph == 0 and pw == 0
    Ho = conv_out_size(H, kh, ph, sh, dh)
    Wo = conv_out_size(W, kw, pw, sw, dw)
    y = np.empty((N, Co, Ho, Wo), F32)
    wm = w.reshape(Co, Ci * kh * kw)
    for n in range(N):
        if is_1x1:
            col = x[n].reshape(C, H * W)
        else:
            col, _, _ = im2col(x[n], kh, kw, ph, pw, sh, sw, dh, dw)
        out = sgemm(wm, col)
        if b is not None:
            # rank-1 GEMM with the ones vector: out += b * 1^T  (beta = 1)
            out = out + b.reshape(Co, 1).astype(F32)
        y[n] = out.reshape(Co, Ho, Wo)
    return y


def deconvolution(x, w, b, stride, pad, dil):
    """DeconvolutionLayer::Forward_cpu deconv_layer.cpp:25-40 -> backward_cpu_gemm
    base_conv_layer.cpp:282-298 (col = W^T x, then col2im) + bias.
    w: [Cin, Cout, kh, kw]; out = s(in-1) + d(k-1) + 1 - 2p (deconv_layer.cpp:8-22)."""
    N, C, H, W = x.shape
    Ci, Co, kh, kw = w.shape
    assert Ci == C
    Ho = stride * (H - 1) + dil * (kh - 1) + 1 - 2 * pad
    Wo = stride * (W - 1) + dil * (kw - 1) + 1 - 2 * pad
    y = np.empty((N, Co, Ho, Wo), F32)
    wt = np.ascontiguousarray(w.reshape(Ci, Co * kh * kw).T)
    for n in range(N):
        col = sgemm(wt, x[n].reshape(C, H * W))
        out = col2im(col, Co, Ho, Wo, kh, kw, pad, pad, stride, stride, dil, dil)
        if b is not None:
            out = out + b.reshape(Co, 1, 1).astype(F32)
        y[n] = out
    return y


def batch_norm_global(x, mean_sum, var_sum, scale_factor, eps=1e-5):
    """BatchNormLayer::Forward_cpu inference branch, batch_norm_layer.cpp:86-93
    (stats * 1/scale_factor, 0 if the factor is 0), :105-111 (subtract mean),
    :137-149 (sqrt(var+eps) via powx 0.5, divide)."""
    sf = F32(0) if scale_factor == 0 else F32(1) / F32(scale_factor)
    mean = (mean_sum.astype(F32) * sf).astype(F32)
    var = (var_sum.astype(F32) * sf).astype(F32)
    t = x - mean.reshape(1, -1, 1, 1)
    std = np.power(var + F32(eps), F32(0.5)).astype(F32)
    return (t / std.reshape(1, -1, 1, 1)).astype(F32)


def scale_bias(x, gamma, beta):
    """ScaleLayer::Forward_cpu scale_layer.cpp:120-133 (x*gamma per channel) then
    BiasLayer::Forward_cpu bias_layer.cpp:72-88 (+beta) as separate roundings."""
    y = (x * gamma.reshape(1, -1, 1, 1).astype(F32)).astype(F32)
    if beta is not None:
        y = (y + beta.reshape(1, -1, 1, 1).astype(F32)).astype(F32)
    return y


def relu(x, negative_slope=0.0):
    """ReLULayer::Forward_cpu relu_layer.cpp:9-19."""
    return (np.maximum(x, F32(0)) + F32(negative_slope) * np.minimum(x, F32(0))).astype(F32)


def eltwise_sum(bottoms, coeffs=None):
    """EltwiseLayer::Forward_cpu SUM, eltwise_layer.cpp:59-65: top=0; top += c_i*b_i."""
    top = np.zeros_like(bottoms[0], F32)
    for i, b in enumerate(bottoms):
        c = F32(1) if not coeffs else F32(coeffs[i])
        top = (top + c * b).astype(F32)
    return top


def pool_out_size(size, k, pad, stride):
    """PoolingLayer::Reshape pooling_layer.cpp:90-107 (ceil mode + pad clip)."""
    out = int(math.ceil(float(size + 2 * pad - k) / stride)) + 1
    if pad and (out - 1) * stride >= size + pad:
        out -= 1
    return out


def max_pool(x, k, stride, pad=0):
    """PoolingLayer::Forward_cpu MAX, pooling_layer.cpp:140-187: windows
    [ph*s-pad, min(+k, H)) clipped at 0, init -FLT_MAX."""
    N, C, H, W = x.shape
    Ho = pool_out_size(H, k, pad, stride)
    Wo = pool_out_size(W, k, pad, stride)
    y = np.full((N, C, Ho, Wo), -np.finfo(F32).max, F32)
    for dy in range(k):
        for dx in range(k):
            # input index for output (ph,pw): ph*s - pad + dy, valid if inside the image
            ph = np.arange(Ho)
            pw = np.arange(Wo)
            iy = ph * stride - pad + dy
            ix = pw * stride - pad + dx
            vy = (iy >= 0) & (iy < H)
            vx = (ix >= 0) & (ix < W)
            if not vy.any() or not vx.any():
                continue
            sub = x[:, :, iy[vy]][:, :, :, ix[vx]]
            ys = np.nonzero(vy)[0]
            xs = np.nonzero(vx)[0]
            view = y[:, :, ys[0]:ys[-1] + 1, xs[0]:xs[-1] + 1]
            np.maximum(view, sub, out=view)
    return y


def crop(a, ref, off_h=0, off_w=0):
    """DeepCut CropLayer crop_layer.cpp:24-50: requires H0-off > H1 strictly."""
    assert a.shape[2] - off_h > ref.shape[2] and a.shape[3] - off_w > ref.shape[3], "invalid offset"
    return np.ascontiguousarray(a[:, :, off_h:off_h + ref.shape[2], off_w:off_w + ref.shape[3]])


def sigmoid(x):
    """SigmoidLayer sigmoid_layer.cpp:9-22: 1/(1+exp(-x)) in fp32."""
    return (F32(1) / (F32(1) + np.exp(-x.astype(F32)))).astype(F32)


def pose_from_mats(prob, loc, scale=1.0, stride=8.0, locref_scale=math.sqrt(53.0)):
    """The demo's read-out, python/pose/estimate_pose.py:131-143 (_pose_from_mats) on the blob layouts of
    :224-243 (_cnn_process_image): prob [14,H,W], loc [28,H,W] -> pose [5,14]."""
    J = prob.shape[0]
    offmat = loc.reshape(J, 2, loc.shape[1], loc.shape[2]).transpose(2, 3, 0, 1)   # [y][x][joint][2]
    scoremat = prob.transpose(1, 2, 0)
    pose = []
    for j in range(J):
        maxloc = np.unravel_index(np.argmax(scoremat[:, :, j]), scoremat[:, :, j].shape)
        offset = np.array(offmat[maxloc][j])[::-1]
        pos_f8 = np.array(maxloc).astype("float") * stride + 0.5 * stride + offset * locref_scale
        pose.append(np.hstack((pos_f8[::-1] / scale, [scoremat[maxloc][j]], offset * locref_scale / scale)))
    return np.array(pose).T


# --------------------------------------------------------------------------
# Net (net.cpp:40-284 Init, :565-581 ForwardFromTo)
# --------------------------------------------------------------------------
def _rep(param, name, default):
    v = param.get(name)
    return v[0] if v else default


class RefNet:
    """Blob-table interpreter of a deploy prototxt, following Net::Init's
    bottom/top wiring by blob NAME (in-place layers overwrite, net.cpp:385-440;
    auto-inserted Split layers only alias data, split_layer.cpp:26-31, so they
    are omitted).  Outputs = blobs never consumed (net.cpp:268-274)."""

    def __init__(self, net_param):
        self.param = net_param
        self.layers = [l for l in net_param.get("layer", [])
                       if not self._filtered(l)]
        self.inputs = list(net_param.get("input", []))
        dims = net_param.get("input_dim", [])
        self.input_shapes = {n: tuple(dims[4 * i:4 * i + 4]) for i, n in enumerate(self.inputs)}
        for sh in net_param.get("input_shape", []):
            self.input_shapes[self.inputs[len(self.input_shapes) - 1]] = tuple(sh["dim"])
        self.params = {}
        self._init_params()

    @staticmethod
    def _filtered(layer):
        # FilterNet, net.cpp:287-310: the deploy net has no include/exclude rules
        for rule in layer.get("include", []):
            if pt.get(rule, "phase") == "TRAIN":
                return True
        return False

    # parameter blob shapes in the reference's orders
    def _init_params(self):
        shapes = {n: s for n, s in self.input_shapes.items()}
        self.param_shapes = {}
        self.layer_types = {}
        for l in self.layers:
            self.layer_types[pt.get(l, "name")] = pt.get(l, "type")
            t = pt.get(l, "type")
            name = pt.get(l, "name")
            bots = l.get("bottom", [])
            tops = l.get("top", [])
            bs = [shapes[b] for b in bots]
            if t in ("Convolution", "Deconvolution"):
                cp = pt.get(l, "convolution_param")
                co = pt.get(cp, "num_output")
                k = _rep(cp, "kernel_size", 1)
                s = _rep(cp, "stride", 1)
                p = _rep(cp, "pad", 0)
                d = _rep(cp, "dilation", 1)
                bias = pt.get(cp, "bias_term", True)
                n, c, h, w = bs[0]
                if t == "Convolution":
                    ws = (co, c, k, k)
                    oh, ow = conv_out_size(h, k, p, s, d), conv_out_size(w, k, p, s, d)
                else:
                    ws = (c, co, k, k)   # reverse_dimensions, base_conv_layer.cpp:125-140
                    oh = s * (h - 1) + d * (k - 1) + 1 - 2 * p
                    ow = s * (w - 1) + d * (k - 1) + 1 - 2 * p
                self.param_shapes[name] = [ws] + ([(co,)] if bias else [])
                shapes[tops[0]] = (n, co, oh, ow)
            elif t == "BatchNorm":
                c = bs[0][1]
                self.param_shapes[name] = [(c,), (c,), (1,)]
                shapes[tops[0]] = bs[0]
            elif t == "Scale":
                c = bs[0][1]
                sp = pt.get(l, "scale_param", {})
                self.param_shapes[name] = [(c,)] + ([(c,)] if pt.get(sp, "bias_term", False) else [])
                shapes[tops[0]] = bs[0]
            elif t == "Pooling":
                pp = pt.get(l, "pooling_param")
                k, s, p = pt.get(pp, "kernel_size"), pt.get(pp, "stride", 1), pt.get(pp, "pad", 0)
                n, c, h, w = bs[0]
                shapes[tops[0]] = (n, c, pool_out_size(h, k, p, s), pool_out_size(w, k, p, s))
            elif t == "Crop":
                shapes[tops[0]] = bs[0][:2] + bs[1][2:]
            elif t in ("ReLU", "Sigmoid", "Eltwise"):
                shapes[tops[0]] = bs[0]
            else:
                raise NotImplementedError("oracle: layer type %s" % t)
        self.blob_shapes = shapes

    def typed_param_shapes(self):
        """{layer: (type, [shapes])} in prototxt order -- the input of the synthetic
        weight recipe (deepcut-cnn_b200/synth.py)."""
        return {n: (self.layer_types[n], s) for n, s in self.param_shapes.items()}

    def reshape_input(self, name, shape):
        self.input_shapes[name] = tuple(shape)
        self._init_params()

    def output_names(self):
        consumed, produced = set(), []
        for l in self.layers:
            consumed.update(l.get("bottom", []))
        for n in self.inputs:
            produced.append(n)
        for l in self.layers:
            for t in l.get("top", []):
                if t not in produced:
                    produced.append(t)
        return sorted(n for n in produced if n not in consumed)

    def forward(self, inputs, want=None):
        """inputs: {name: ndarray NCHW}; returns {output blob name: ndarray}
        (+ any intermediate named in ``want``, captured right after the layer of
        that name runs, like Net::ForwardFromTo with debug_info)."""
        blobs = {k: np.ascontiguousarray(v, F32) for k, v in inputs.items()}
        keep = {}
        P = self.params
        for l in self.layers:
            t = pt.get(l, "type")
            name = pt.get(l, "name")
            bots = [blobs[b] for b in l.get("bottom", [])]
            top = l.get("top", [])[0]
            if t in ("Convolution", "Deconvolution"):
                cp = pt.get(l, "convolution_param")
                s, p, d = _rep(cp, "stride", 1), _rep(cp, "pad", 0), _rep(cp, "dilation", 1)
                w = P[name][0]
                b = P[name][1] if len(P[name]) > 1 else None
                fn = convolution if t == "Convolution" else deconvolution
                y = fn(bots[0], w, b, s, p, d)
            elif t == "BatchNorm":
                bp = pt.get(l, "batch_norm_param", {})
                assert pt.get(bp, "use_global_stats", True), "oracle covers TEST phase only"
                y = batch_norm_global(bots[0], P[name][0], P[name][1], P[name][2][0],
                                      pt.get(bp, "eps", 1e-5))
            elif t == "Scale":
                y = scale_bias(bots[0], P[name][0], P[name][1] if len(P[name]) > 1 else None)
            elif t == "ReLU":
                y = relu(bots[0], pt.get(pt.get(l, "relu_param", {}), "negative_slope", 0.0))
            elif t == "Eltwise":
                ep = pt.get(l, "eltwise_param", {})
                assert pt.get(ep, "operation", "SUM") == "SUM"
                y = eltwise_sum(bots, ep.get("coeff"))
            elif t == "Pooling":
                pp = pt.get(l, "pooling_param")
                assert pt.get(pp, "pool", "MAX") == "MAX"
                y = max_pool(bots[0], pt.get(pp, "kernel_size"), pt.get(pp, "stride", 1),
                             pt.get(pp, "pad", 0))
            elif t == "Crop":
                cp = pt.get(l, "crop_param", {})
                y = crop(bots[0], bots[1], pt.get(cp, "offset_height", 0), pt.get(cp, "offset_width", 0))
            elif t == "Sigmoid":
                y = sigmoid(bots[0])
            else:
                raise NotImplementedError(t)
            blobs[top] = y
            if want and name in want:
                keep[name] = y
        out = {n: blobs[n] for n in self.output_names()}
        out.update(keep)
        return out


def load_net(path):
    return RefNet(pt.parse_file(path))
