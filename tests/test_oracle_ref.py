"""Pins the numpy restatement (oracle/caffe_ref.py) to THE REFERENCE ITSELF: its CPU layer sources compiled
verbatim from /root/reference into oracle/_ref/librefcaffe.so (oracle/build_ref.py).  Every layer type on the
DeeperCut path is run through the reference's LayerSetUp / Reshape / Forward_cpu on the same seeded inputs, in
the geometries the path uses plus the edge cases the reference's tests cover (ragged sizes, ceil-mode pooling
overhang, dilation + padding, crop offsets).  Skipped only where neither /root/reference nor a prebuilt
library exists."""
import numpy as np
import pytest

import dcutil
import netutil
from oracle import caffe_ref as cr

pytestmark = pytest.mark.skipif(not netutil.reference_available(), reason="oracle/_ref not built and /root/reference absent")

TOL = 2e-5      # fp32 GEMM summation order (OpenBLAS blocks vs numpy) on O(1) values


def run_ref(text, inputs, params=None, want=None):
    from oracle import ref_caffe
    net = ref_caffe.RefCaffeNet(text)
    if params:
        net.set_params(params)
    return net.forward(inputs, want=want)


def header(shapes):
    out = ['name: "t"']
    for name, s in shapes.items():
        out.append('input: "%s"' % name)
        out += ["input_dim: %d" % d for d in s]
    return "\n".join(out) + "\n"


@pytest.mark.parametrize("cin,cout,k,stride,pad,dil,hw,bias", [
    (3, 8, 7, 2, 3, 1, (37, 50), True),         # conv1 geometry, ragged size
    (16, 12, 1, 1, 0, 1, (9, 13), False),       # 1x1
    (16, 8, 1, 2, 0, 1, (9, 13), False),        # strided 1x1 projection (res3a_branch1)
    (8, 8, 3, 1, 1, 1, (10, 7), False),         # 3x3
    (8, 10, 3, 1, 2, 2, (11, 12), True),        # dilated res5 3x3
    (4, 6, 3, 2, 4, 3, (15, 9), True),          # stride + dilation + large pad
])
def test_convolution_matches_reference(cin, cout, k, stride, pad, dil, hw, bias):
    rng = np.random.default_rng(cin * 131 + k)
    x = rng.standard_normal((2, cin) + hw).astype(np.float32)
    w = rng.standard_normal((cout, cin, k, k)).astype(np.float32) * 0.2
    b = rng.standard_normal(cout).astype(np.float32) if bias else None
    text = header({"x": x.shape}) + '''layer { name: "c" type: "Convolution" bottom: "x" top: "y"
      convolution_param { num_output: %d kernel_size: %d stride: %d pad: %d dilation: %d bias_term: %s } }''' % (
        cout, k, stride, pad, dil, "true" if bias else "false")
    got = run_ref(text, {"x": x}, {"c": [w] + ([b] if bias else [])})["y"]
    want = cr.convolution(x, w, b, stride, pad, dil)
    assert got.shape == want.shape
    assert netutil.max_err(got, want) < TOL


@pytest.mark.parametrize("cin,cout,k,stride,pad,hw", [(16, 14, 3, 2, 0, (8, 10)), (8, 5, 4, 2, 1, (5, 7)), (6, 6, 3, 1, 1, (6, 6))])
def test_deconvolution_matches_reference(cin, cout, k, stride, pad, hw):
    rng = np.random.default_rng(cout)
    x = rng.standard_normal((2, cin) + hw).astype(np.float32)
    w = rng.standard_normal((cin, cout, k, k)).astype(np.float32) * 0.2
    b = rng.standard_normal(cout).astype(np.float32)
    text = header({"x": x.shape}) + '''layer { name: "d" type: "Deconvolution" bottom: "x" top: "y"
      convolution_param { num_output: %d kernel_size: %d stride: %d pad: %d } }''' % (cout, k, stride, pad)
    got = run_ref(text, {"x": x}, {"d": [w, b]})["y"]
    want = cr.deconvolution(x, w, b, stride, pad, 1)
    assert got.shape == want.shape
    assert netutil.max_err(got, want) < TOL


def test_batchnorm_scale_relu_chain_matches_reference():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 12, 7, 9)).astype(np.float32) * 3
    mean, var = rng.standard_normal(12).astype(np.float32) * 4, rng.uniform(0.5, 9, 12).astype(np.float32)
    sf = np.array([4.0], np.float32)                 # moving-average factor: stats are stored multiplied by it
    gamma, beta = rng.uniform(0.5, 2, 12).astype(np.float32), rng.standard_normal(12).astype(np.float32)
    text = header({"x": x.shape}) + '''
      layer { name: "bn" type: "BatchNorm" bottom: "x" top: "x" batch_norm_param { use_global_stats: true } }
      layer { name: "sc" type: "Scale" bottom: "x" top: "x" scale_param { bias_term: true } }
      layer { name: "re" type: "ReLU" bottom: "x" top: "x" }'''
    got = run_ref(text, {"x": x}, {"bn": [mean * 4, var * 4, sf], "sc": [gamma, beta]})["x"]
    want = cr.relu(cr.scale_bias(cr.batch_norm_global(x, mean * 4, var * 4, 4.0), gamma, beta))
    assert netutil.max_err(got, want) < TOL
    # scale_factor 0 means "no statistics yet": the reference divides by 1 (batch_norm_layer.cpp:98-100)
    got0 = run_ref(text, {"x": x}, {"bn": [mean, var, np.zeros(1, np.float32)], "sc": [gamma, beta]})["x"]
    want0 = cr.relu(cr.scale_bias(cr.batch_norm_global(x, mean, var, 0.0), gamma, beta))
    assert netutil.max_err(got0, want0) < TOL


@pytest.mark.parametrize("k,stride,pad,hw", [(3, 2, 0, (112, 112)), (3, 2, 0, (17, 22)), (2, 2, 0, (7, 9)), (3, 2, 1, (9, 9)), (3, 1, 1, (5, 6))])
def test_max_pool_ceil_mode_matches_reference(k, stride, pad, hw):
    x = np.random.default_rng(k + hw[0]).standard_normal((2, 5) + hw).astype(np.float32)
    text = header({"x": x.shape}) + '''layer { name: "p" type: "Pooling" bottom: "x" top: "y"
      pooling_param { pool: MAX kernel_size: %d stride: %d pad: %d } }''' % (k, stride, pad)
    got = run_ref(text, {"x": x})["y"]
    want = cr.max_pool(x, k, stride, pad)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_eltwise_crop_sigmoid_match_reference():
    rng = np.random.default_rng(9)
    a, b = rng.standard_normal((2, 6, 9, 11)).astype(np.float32), rng.standard_normal((2, 6, 9, 11)).astype(np.float32)
    big, small = rng.standard_normal((2, 4, 12, 15)).astype(np.float32), np.zeros((2, 7, 9, 11), np.float32)
    text = header({"a": a.shape, "b": b.shape, "big": big.shape, "small": small.shape}) + '''
      layer { name: "e" type: "Eltwise" bottom: "a" bottom: "b" top: "sum" }
      layer { name: "e2" type: "Eltwise" bottom: "a" bottom: "b" top: "wsum" eltwise_param { operation: SUM coeff: 0.5 coeff: -2 } }
      layer { name: "c0" type: "Crop" bottom: "big" bottom: "small" top: "c0" }
      layer { name: "c1" type: "Crop" bottom: "big" bottom: "small" top: "c1" crop_param { offset_height: 2 offset_width: 3 } }
      layer { name: "s" type: "Sigmoid" bottom: "sum" top: "sig" }'''
    got = run_ref(text, {"a": a, "b": b, "big": big, "small": small})
    assert np.array_equal(got["sum"], cr.eltwise_sum([a, b]))
    assert netutil.max_err(got["wsum"], cr.eltwise_sum([a, b], [0.5, -2.0])) < 1e-6
    assert np.array_equal(got["c0"], cr.crop(big, small))
    assert np.array_equal(got["c1"], cr.crop(big, small, 2, 3))
    assert netutil.max_err(got["sig"], cr.sigmoid(a + b)) < 1e-6


def test_tiny_net_every_blob_matches_reference(tmp_path):
    """Whole DeeperCut topology (one block per stage), ragged input: every named blob of the reference net."""
    path, weights = netutil.build(tmp_path, (1, 1, 1, 1), 64, 64)
    x = dcutil.synth.images(2, 72, 88, seed=5)
    ref = netutil.reference_forward(path, weights, x)
    finals = {}                                   # blob name -> name of the LAST layer writing it (in-place chains)
    for l in cr.load_net(path).layers:
        finals[l["top"][0]] = l["name"][0] if isinstance(l["name"], list) else l["name"]
    got = netutil.oracle_forward(path, weights, x, want=set(finals.values()))
    checked = 0
    for blob, layer in finals.items():
        assert ref[blob].shape == got[layer].shape, blob
        assert netutil.max_err(ref[blob], got[layer]) < 5e-5, blob
        checked += 1
    assert checked >= 30
    for k in ("prob", "loc_pred", "next_pred"):
        assert netutil.max_err(ref[k], got[k]) < TOL, k


def test_resnet152_structure_matches_reference_net():
    """The generated 680-layer deploy net builds in the reference's own SetUp code with the blob shapes and
    parameter counts the numpy oracle derives (so the synthetic-weight recipe feeds both identically)."""
    import importlib
    gen = importlib.import_module("deepcut-cnn_b200.gen_prototxt")
    from oracle import ref_caffe, prototxt as opt
    text = gen.generate(gen.STAGES_152, 64, 64)
    ref = ref_caffe.RefCaffeNet(text)
    net = cr.RefNet(opt.parse(text))
    real = [(n, t) for n, t in zip(ref.layer_names, ref.layer_types) if t != "Split"]
    assert len(real) == 680
    assert [n for n, _ in real] == [l["name"] if not isinstance(l["name"], list) else l["name"][0] for l in net.layers]
    counts = ref.param_counts()
    want = {n: [int(np.prod(s)) for s in shapes] for n, shapes in net.param_shapes.items()}
    assert counts == want
    for blob, shape in net.blob_shapes.items():
        assert ref._blob(blob).shape == tuple(shape), blob


def test_resnet152_forward_matches_reference(tmp_path):
    """Full ResNet-152 DeeperCut at 96x128 with the calibrated synthetic weights: numpy oracle vs the reference."""
    import importlib
    gen = importlib.import_module("deepcut-cnn_b200.gen_prototxt")
    path = str(tmp_path / "r152.prototxt")
    gen.write(path, height=96, width=128)
    weights = dcutil.synth.calibrated_weights(dcutil.ptx.parse_file(path))
    x = dcutil.synth.images(1, 96, 128, seed=21)
    ref = netutil.reference_forward(path, weights, x, want=["prob", "loc_pred", "next_pred"])
    got = netutil.oracle_forward(path, weights, x)
    for k in ("prob", "loc_pred", "next_pred"):
        assert netutil.max_err(ref[k], got[k]) < 1e-4, k
