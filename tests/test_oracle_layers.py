"""Pins the CPU oracle (oracle/caffe_ref.py) against every known-answer / golden vector the
reference's own tests hold for this path.  Each test cites the reference test
it restates; the expected numbers are the reference's."""
import numpy as np
import pytest

from oracle import caffe_ref as R


def naive_conv(x, w, b, stride, pad, dil):
    """Independent 7-loop convolution in the spirit of caffe_conv
    (src/caffe/test/test_convolution_layer.cpp:19-139), float64 accumulate."""
    (sh, sw), (ph, pw), (dh, dw) = R._hw(stride), R._hw(pad), R._hw(dil)
    N, C, H, W = x.shape
    Co, _, kh, kw = w.shape
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    y = np.zeros((N, Co, Ho, Wo))
    for n in range(N):
        for o in range(Co):
            for yy in range(Ho):
                for xx in range(Wo):
                    acc = 0.0
                    for c in range(C):
                        for p in range(kh):
                            for q in range(kw):
                                iy, ix = yy * sh - ph + p * dh, xx * sw - pw + q * dw
                                if 0 <= iy < H and 0 <= ix < W:
                                    acc += float(x[n, c, iy, ix]) * float(w[o, c, p, q])
                    y[n, o, yy, xx] = acc + (float(b[o]) if b is not None else 0.0)
    return y


@pytest.mark.parametrize("k,stride,pad,dil", [
    (3, 2, 0, 1),      # TestSimpleConvolution        test_convolution_layer.cpp:231-265
    (3, 1, 0, 2),      # TestDilatedConvolution       :267-309  (input enlarged as in the test)
    (1, 1, 0, 1),      # Test1x1Convolution           :443-468
    (3, 1, 2, 2),      # the res5 geometry: pad = dilation = 2
    (7, 2, 3, 1),      # conv1
])
def test_convolution_vs_naive_reference(k, stride, pad, dil):
    rng = np.random.default_rng(1701)
    shape = (2, 3, 8, 7) if dil > 1 else (2, 3, 6, 4) if k < 7 else (1, 3, 15, 13)
    x = rng.standard_normal(shape).astype(np.float32)
    w = rng.standard_normal((4, 3, k, k)).astype(np.float32)
    b = np.full(4, 0.1, np.float32)
    assert np.abs(R.convolution(x, w, b, stride, pad, dil) - naive_conv(x, w, b, stride, pad, dil)).max() < 1e-4


def test_sobel_separable_identity():
    # TestSobelConvolution (:498-590): a 3x3 Sobel x-filter (stride (2,1) h/w) equals the column filter
    # [1 2 1]^T (3x1, stride 2 in h) followed by the row filter [-1 0 1] (1x3).
    rng = np.random.default_rng(1701)
    x = rng.standard_normal((2, 3, 6, 4)).astype(np.float32)
    sob = np.zeros((1, 3, 3, 3), np.float32)
    sob[:, :] = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], np.float32)
    full = R.convolution(x, sob, None, (2, 1), 0, 1)
    colf = np.zeros((1, 3, 3, 1), np.float32)
    colf[:, :, :, 0] = np.array([1, 2, 1], np.float32)
    rowf = np.zeros((1, 1, 1, 3), np.float32)
    rowf[0, 0, 0] = np.array([-1, 0, 1], np.float32)
    sep = R.convolution(R.convolution(x, colf, None, (2, 1), 0, 1), rowf, None, 1, 0, 1)
    assert full.shape == sep.shape and np.abs(full - sep).max() < 1e-4


def test_deconvolution_overlap_known_answer():
    # DeconvolutionLayerTest.TestSimpleDeconvolution (test_deconvolution_layer.cpp:91-137): ones input,
    # ones weights, bias 0.1, k3 s2: 3.1 / 6.1 / 12.1 by overlap.
    x = np.ones((2, 3, 6, 4), np.float32)
    w = np.ones((3, 4, 3, 3), np.float32)
    b = np.full(4, 0.1, np.float32)
    y = R.deconvolution(x, w, b, 2, 0, 1)
    assert y.shape == (2, 4, 13, 9)
    for h in range(13):
        for ww in range(9):
            ho = h % 2 == 0 and 0 < h < 12
            wo = ww % 2 == 0 and 0 < ww < 8
            exp = 3.1 * (2 if ho else 1) * (2 if wo else 1) - 0.1 * ((2 if ho else 1) * (2 if wo else 1) - 1)
            assert np.allclose(y[:, :, h, ww], exp, atol=1e-4), (h, ww)
    assert np.isclose(y[0, 0, 0, 0], 3.1) and np.isclose(y[0, 0, 2, 1], 6.1) and np.isclose(y[0, 0, 2, 2], 12.1)


def test_gemm_known_answers():
    # GemmTest (test_util_blas.cpp:20-88): A = 1..6 (2x3), B = 1..12 (3x4) in all four transpose combos
    A = np.arange(1, 7, dtype=np.float32).reshape(2, 3)
    B = np.arange(1, 13, dtype=np.float32).reshape(3, 4)
    want = np.array([38, 44, 50, 56, 83, 98, 113, 128], np.float32).reshape(2, 4)
    assert np.array_equal(R.sgemm(A, B), want)
    At = np.array([1, 4, 2, 5, 3, 6], np.float32).reshape(3, 2)
    Bt = np.array([1, 5, 9, 2, 6, 10, 3, 7, 11, 4, 8, 12], np.float32).reshape(4, 3)
    assert np.array_equal(R.sgemm(At.T, B), want)
    assert np.array_equal(R.sgemm(At.T, Bt.T), want)
    assert np.array_equal(R.sgemm(A, Bt.T), want)


def test_max_pool_literal_matrices():
    # PoolingLayerTest.TestForwardSquare (test_pooling_layer.cpp:48-110): 2x2 / stride 1 over
    # [1 2 5 2 3; 9 4 1 4 8; 1 2 5 2 3] -> [9 5 5 8; 9 5 5 8]
    img = np.array([[1, 2, 5, 2, 3], [9, 4, 1, 4, 8], [1, 2, 5, 2, 3]], np.float32)
    x = np.broadcast_to(img, (2, 2, 3, 5)).copy()
    y = R.max_pool(x, 2, 1)
    assert y.shape == (2, 2, 2, 4)
    assert np.array_equal(y[1, 1], np.array([[9, 5, 5, 8], [9, 5, 5, 8]], np.float32))
    # ceil mode + clipping at the border: pool1's 3x3/2 on odd sizes (pooling_layer.cpp:90-93)
    x = np.arange(36, dtype=np.float32).reshape(1, 1, 6, 6)
    y = R.max_pool(x, 3, 2)
    assert y.shape == (1, 1, 3, 3)          # ceil((6-3)/2)+1 = 3: the last window is clipped
    assert np.array_equal(y[0, 0], np.array([[14, 16, 17], [26, 28, 29], [32, 34, 35]], np.float32))
    # padded case shape rule (test_pooling_layer.cpp:478-521 geometry: 3x3 pad 1 stride 2 on 3x3 -> 2x2)
    assert R.pool_out_size(3, 3, 1, 2) == 2
    # PoolingLayerTest.TestForwardMaxPadded (test_pooling_layer.cpp:478-521): 3x3 / stride 2 / pad 2 over
    # [1 2 4; 2 3 2; 4 2 1] -> 3x3 [1 4 4; 4 4 4; 4 4 1] (windows clipped to the image, exact)
    x = np.array([[1, 2, 4], [2, 3, 2], [4, 2, 1]], np.float32).reshape(1, 1, 3, 3)
    y = R.max_pool(x, 3, 2, pad=2)
    assert y.shape == (1, 1, 3, 3)
    assert np.array_equal(y[0, 0], np.array([[1, 4, 4], [4, 4, 4], [4, 4, 1]], np.float32))


def test_gemv_known_answers_and_bias_broadcast():
    # GemmTest.TestGemvCPUGPU (test_util_blas.cpp:89-130): A = [1 2 3; 4 5 6], A x = [14 32], A^T [1 2] = [9 12 15] -- the
    # K = 1 / N = 1 GEMMs the reference's bias path is made of (base_conv_layer.cpp:274-280), exact integers
    A = np.array([[1, 2, 3], [4, 5, 6]], np.float32)
    assert np.array_equal(R.sgemm(A, np.array([[1], [2], [3]], np.float32))[:, 0], np.array([14, 32], np.float32))
    assert np.array_equal(R.sgemm(A.T.copy(), np.array([[1], [2]], np.float32))[:, 0], np.array([9, 12, 15], np.float32))
    # forward_cpu_bias: Y += bias * ones^T as a K = 1 GEMM == a per-channel broadcast add
    # (BiasLayerTest.TestForwardBroadcastMiddle, test_bias_layer.cpp:199-316: y = x + b broadcast over axis 1), 1e-5
    rng = np.random.default_rng(99)
    x = rng.standard_normal((2, 3, 4, 5)).astype(np.float32)
    b = rng.standard_normal(3).astype(np.float32)
    y = R.scale_bias(x, np.ones(3, np.float32), b)
    for c in range(3):
        assert np.abs(y[:, c] - (x[:, c] + b[c])).max() < 1e-5
    w = rng.standard_normal((3, 3, 1, 1)).astype(np.float32)
    yc = R.convolution(x, w, b, 1, 0, 1)                      # the head layers' bias term through the conv path itself
    ref = np.einsum("oi,nihw->nohw", w[:, :, 0, 0].astype(np.float64), x.astype(np.float64)) + b.reshape(1, 3, 1, 1)
    assert np.abs(yc - ref).max() < 1e-5


def test_scale_bias_eltwise_relu_sigmoid_properties():
    rng = np.random.default_rng(1701)
    x = rng.standard_normal((2, 3, 4, 5)).astype(np.float32)
    g = rng.standard_normal(3).astype(np.float32)
    b = rng.standard_normal(3).astype(np.float32)
    y = R.scale_bias(x, g, b)            # ScaleLayerTest.TestForwardBroadcastMiddleWithParamAndBias (test_scale_layer.cpp:317-342), 1e-5
    for c in range(3):
        assert np.abs(y[:, c] - (x[:, c] * g[c] + b[c])).max() < 1e-5
    s = R.eltwise_sum([x, 2 * x, x])     # EltwiseLayerTest.TestSum (test_eltwise_layer.cpp:87-104), 1e-4
    assert np.abs(s - 4 * x).max() < 1e-4
    assert np.abs(R.eltwise_sum([x, x], [1.0, -0.5]) - 0.5 * x).max() < 1e-5     # TestSumCoeff
    r = R.relu(x)                        # NeuronLayerTest.TestReLU (test_neuron_layer.cpp:208-221)
    assert (r >= 0).all() and np.array_equal(r[x > 0], x[x > 0]) and not r[x <= 0].any()
    rl = R.relu(x, 0.01)                 # TestReLUWithNegativeSlope
    assert np.allclose(rl[x < 0], 0.01 * x[x < 0])
    sg = R.sigmoid(x)                    # TestSigmoid (:321-336)
    assert np.abs(sg - 1.0 / (1.0 + np.exp(-x.astype(np.float64)))).max() < 1e-6 and ((sg >= 0) & (sg <= 1)).all()


def test_im2col_matches_loops_with_dilation():
    # Im2colKernelTest geometry (test_im2col_kernel.cu:55-58: dilation 3, stride 2, pad 1): bit-exact
    rng = np.random.default_rng(1701)
    x = rng.standard_normal((4, 9, 11)).astype(np.float32)
    kh = kw = 3
    col, ho, wo = R.im2col(x, kh, kw, 1, 1, 2, 2, 3, 3)
    assert col.shape == (4 * 9, ho * wo)
    for c in range(4):
        for p in range(kh):
            for q in range(kw):
                for oy in range(ho):
                    for ox in range(wo):
                        iy, ix = -1 + p * 3 + oy * 2, -1 + q * 3 + ox * 2
                        want = x[c, iy, ix] if (0 <= iy < 9 and 0 <= ix < 11) else 0.0
                        assert col[(c * kh + p) * kw + q, oy * wo + ox] == want
    # col2im is its adjoint: <im2col(x), c> == <x, col2im(c)>
    cvec = rng.standard_normal(col.shape).astype(np.float32)
    back = R.col2im(cvec, 4, 9, 11, 3, 3, 1, 1, 2, 2, 3, 3)
    assert abs(float((col.astype(np.float64) * cvec).sum()) - float((x.astype(np.float64) * back).sum())) < 1e-3


def test_batch_norm_global_stats_and_crop_unpinned_rows():
    # The reference has NO test for BatchNorm use_global_stats nor for its CropLayer:
    # check the restatement against the formula in float64 and plain slicing.
    rng = np.random.default_rng(1701)
    x = rng.standard_normal((2, 5, 3, 4)).astype(np.float32) * 3
    mean = rng.standard_normal(5).astype(np.float32)
    var = rng.uniform(0.5, 2, 5).astype(np.float32)
    for factor in (1.0, 4.0):
        y = R.batch_norm_global(x, mean, var, factor)
        want = (x.astype(np.float64) - mean.reshape(1, -1, 1, 1) / factor) / np.sqrt(var.reshape(1, -1, 1, 1) / factor + 1e-5)
        assert np.abs(y - want).max() < 1e-5
    y0 = R.batch_norm_global(x, mean, var, 0.0)       # scale_factor == 0 -> statistics read as 0 (batch_norm_layer.cpp:88-89)
    assert np.allclose(y0, x / np.sqrt(np.float32(1e-5)), rtol=1e-5)
    a = rng.standard_normal((1, 2, 9, 7)).astype(np.float32)
    ref = np.zeros((1, 2, 8, 6), np.float32)
    assert np.array_equal(R.crop(a, ref), a[:, :, :8, :6])
    assert np.array_equal(R.crop(a, np.zeros((1, 2, 4, 3), np.float32), 2, 1), a[:, :, 2:6, 1:4])
    with pytest.raises(AssertionError):                # CHECK_GT is strict (crop_layer.cpp:28-31)
        R.crop(a, np.zeros((1, 2, 9, 7), np.float32))
