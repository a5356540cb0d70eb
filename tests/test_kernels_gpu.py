"""Kernel parity (GPU): every CUDA kernel through the C ABI vs the CPU oracle on seeded inputs.
Tolerances: the reference's own layer tests use 1e-4 abs on O(1) data
(src/caffe/test/test_convolution_layer.cpp:256); integer/copy kernels are bit-exact."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import dcutil
from dcutil import libdc
from oracle import caffe_ref


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import gpuharness
    gpuharness.init()
    return gpuharness


def _bn_params(rng, c):
    return (rng.uniform(0.5, 1.5, c).astype(np.float32), rng.normal(0, 0.3, c).astype(np.float32))


# (cin, cout, k, pad, dil, H, W, N): one per distinct stride-1 backbone shape class of SURVEY 8(a'),
# at spatial sizes with ragged tile edges (43x43 is the prototxt's own res5 size at 688x688).
CONV_CASES = [
    (64, 64, 1, 0, 1, 19, 23, 2),
    (64, 64, 3, 1, 1, 19, 23, 2),
    (64, 256, 1, 0, 1, 16, 16, 1),
    (256, 64, 1, 0, 1, 9, 31, 2),
    (128, 128, 3, 1, 1, 33, 17, 1),
    (256, 256, 3, 1, 1, 16, 16, 2),
    (1024, 256, 1, 0, 1, 8, 13, 1),
    (256, 1024, 1, 0, 1, 8, 13, 2),
    (512, 512, 3, 2, 2, 43, 43, 1),       # res5 dilated conv (test_convolution_layer.cpp:267 analogue)
    (512, 2048, 1, 0, 1, 11, 9, 1),
    (2048, 512, 1, 0, 1, 11, 9, 1),
    (64, 64, 3, 3, 3, 21, 18, 1),         # dilation 3 (test_im2col_kernel.cu:55-58 fixture)
    # more work units than SMs with a short last wave (150 / 198 units on 148 SMs); 192 outputs: ragged second channel tile
    (128, 128, 1, 0, 1, 120, 160, 1),
    (64, 128, 1, 0, 1, 132, 192, 1),
    (64, 192, 3, 1, 1, 40, 48, 5),
    # long-K 1x1 reduce convs with enough pixels for the 256-channel-tile CTA-pair kernel (conv_igemm<256, 2, 8>, single-buffered
    # accumulators): res4 branch2a (one 256-channel tile) and res5 branch2a (two), 45x80 maps with a ragged last pixel tile
    (1024, 256, 1, 0, 1, 45, 80, 6),
    (2048, 512, 1, 0, 1, 45, 83, 3),
    # 3x3 convs with at least as many pixel tiles as SMs: the CTA-pair form of the 128-channel-tile 3x3 kernel (two vertically
    # adjacent 16x8 tiles per pair, 45 rows = ragged last pair, odd tile counts = a phantom tile in the last pair)
    (256, 256, 3, 1, 1, 45, 80, 6),
    (512, 512, 3, 2, 2, 45, 80, 5),
    (128, 128, 3, 1, 1, 37, 83, 9),
    # 64 -> 64 channel 3x3 convs with at least as many 16x8 tiles as SMs: TALL mode (one tall box per column offset, resident weights),
    # ragged in both directions, dilation 1 and 2, an odd tile count (phantom tile in the last pair)
    (64, 64, 3, 1, 1, 45, 83, 6),
    (64, 64, 3, 2, 2, 41, 70, 7),
    (64, 64, 3, 1, 1, 72, 100, 3),          # 7 x 9 x 3 = 189 tiles: odd
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_bn_relu_matches_oracle(case, _gpu):
    ci, co, k, pad, dil, h, w, n = case
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    x = rng.standard_normal((n, ci, h, w)).astype(np.float32)
    x *= (rng.random((n, ci, h, w)) > 0.3)            # post-ReLU-like sparsity
    wt = (rng.standard_normal((co, ci, k, k)) * np.sqrt(2.0 / (ci * k * k))).astype(np.float32)
    a, b = _bn_params(rng, co)
    ref = caffe_ref.convolution(x, wt, None, 1, pad, dil)
    ref = np.maximum(ref * a.reshape(1, -1, 1, 1) + b.reshape(1, -1, 1, 1), 0)
    got = _gpu.conv_bn(x, wt, a, b, pad=pad, dil=dil, relu=True)
    assert got.shape == ref.shape
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() < 1e-4, np.abs(got - ref).max()


@pytest.mark.parametrize("case", [(1024, 256, 1, 0, 1, 45, 80, 6), (2048, 512, 1, 0, 1, 45, 83, 3), (256, 256, 3, 1, 1, 45, 80, 6), (512, 512, 3, 2, 2, 45, 80, 5),
                                  (64, 64, 3, 1, 1, 45, 83, 6), (64, 64, 3, 2, 2, 41, 70, 7)])
def test_tile_and_pairing_choices_are_bitwise_neutral(case, _gpu, monkeypatch):
    """dc_conv_forward's kernel choices (CTA pairs or single CTAs for the 3x3 convs, 128- or 256-channel tiles, TALL mode or one box
    per tap for the 64 -> 64 channel convs; off-by-default forms included) keep every output element's K chain: the same launch under
    each switch gives bitwise the same tensor."""
    ci, co, k, pad, dil, h, w, n = case
    rng = np.random.default_rng(ci + co + k)
    x = np.maximum(rng.standard_normal((n, ci, h, w)), 0).astype(np.float32)
    wt = (rng.standard_normal((co, ci, k, k)) * np.sqrt(2.0 / (ci * k * k))).astype(np.float32)
    a, b = _bn_params(rng, co)
    base = _gpu.conv_bn(x, wt, a, b, pad=pad, dil=dil, relu=True)
    for var, val in (("DC_CONV_BN256", "2"), ("DC_CONV_PAIR_3X3", "0"), ("DC_CONV_PAIR_ALL", "1"), ("DC_CONV_TALL", "0")):
        monkeypatch.setenv(var, val)
        got = _gpu.conv_bn(x, wt, a, b, pad=pad, dil=dil, relu=True)
        monkeypatch.delenv(var)
        assert np.array_equal(got, base), (var, float(np.abs(got - base).max()))


def test_conv_residual_add_relu(_gpu):
    rng = np.random.default_rng(7)
    n, ci, co, h, w = 2, 256, 1024, 12, 10
    x = rng.standard_normal((n, ci, h, w)).astype(np.float32)
    wt = (rng.standard_normal((co, ci, 1, 1)) * np.sqrt(2.0 / ci)).astype(np.float32)
    a, b = _bn_params(rng, co)
    shortcut = rng.standard_normal((n, co, h, w)).astype(np.float32)
    branch = caffe_ref.convolution(x, wt, None, 1, 0, 1) * a.reshape(1, -1, 1, 1) + b.reshape(1, -1, 1, 1)
    ref = caffe_ref.relu(caffe_ref.eltwise_sum([shortcut, branch.astype(np.float32)]))
    got = _gpu.conv_bn(x, wt, a, b, relu=True, residual_nchw=shortcut)
    assert np.abs(got - ref).max() < 1e-4
    # no-ReLU / no-residual variant (projection shortcut branch1: conv + BN + Scale only)
    got2 = _gpu.conv_bn(x, wt, a, b, relu=False)
    assert np.abs(got2 - branch).max() < 1e-4 and (got2 < 0).any()


# (cin, cout, k, pad, dil, H, W, N, residual): few work units, long K loops -> dc_conv_forward shares each unit's K loop
# among a cluster of 4 / 2 CTAs.  res4 / res5 / res3 geometry of one 512x512 image and smaller; ragged tiles included.
SPLIT_K_CASES = [
    (256, 256, 3, 1, 1, 32, 32, 1, False),      # res4 2b @512^2: 8 pixel tiles x 4 channel tiles, 36 K-steps -> S = 4
    (1024, 256, 1, 0, 1, 32, 32, 1, False),     # res4 2a: 16 K-steps -> S = 4
    (512, 512, 3, 2, 2, 32, 32, 1, False),      # res5 2b dilated: 64 units -> S = 2
    (2048, 512, 1, 0, 1, 19, 13, 1, True),      # ragged pixel tile + residual + ReLU
    (128, 128, 3, 1, 1, 21, 17, 2, True),       # 3x3, ragged rectangles, residual
    (1024, 160, 1, 0, 1, 9, 11, 1, True),       # ragged channel tile (160 of 256 packed rows), 16 K-steps
]


@pytest.mark.parametrize("case", SPLIT_K_CASES)
def test_split_k_matches_oracle_and_unsplit_kernel(case, _gpu):
    ci, co, k, pad, dil, h, w, n, with_res = case
    L = libdc.lib()
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    x = np.maximum(rng.standard_normal((n, ci, h, w)), 0).astype(np.float32)
    wt = (rng.standard_normal((co, ci, k, k)) * np.sqrt(2.0 / (ci * k * k))).astype(np.float32)
    a, b = _bn_params(rng, co)
    shortcut = rng.standard_normal((n, co, h, w)).astype(np.float32) if with_res else None
    ref = caffe_ref.convolution(x, wt, None, 1, pad, dil) * a.reshape(1, -1, 1, 1) + b.reshape(1, -1, 1, 1)
    if with_res:
        ref = ref + shortcut
    ref = np.maximum(ref, 0)
    got = {}
    try:
        for s in (1, 2, 4):
            libdc.check(L.dc_set_split_k(s))
            got[s] = _gpu.conv_bn(x, wt, a, b, pad=pad, dil=dil, relu=True, residual_nchw=shortcut)
            assert np.isfinite(got[s]).all()
            assert np.abs(got[s] - ref).max() < 1e-4, (s, np.abs(got[s] - ref).max())
            again = _gpu.conv_bn(x, wt, a, b, pad=pad, dil=dil, relu=True, residual_nchw=shortcut)
            assert np.array_equal(got[s], again), "split %d is not deterministic" % s
        rows = _gpu.conv_bn(x, wt, a, b, pad=pad, dil=dil, relu=False, f32_rows=True)       # fp32-rows epilogue, split
        libdc.check(L.dc_set_split_k(1))
        rows1 = _gpu.conv_bn(x, wt, a, b, pad=pad, dil=dil, relu=False, f32_rows=True)
    finally:
        libdc.check(L.dc_set_split_k(4))
    # the split only re-associates the fp32 sum over K
    for s in (2, 4):
        assert np.abs(got[s] - got[1]).max() < 5e-5, (s, np.abs(got[s] - got[1]).max())
    assert np.abs(rows[:, :co] - rows1[:, :co]).max() < 5e-5
    assert L.dc_set_split_k(3) != 0 and L.dc_get_split_k() == 4
    assert L.dc_set_split_k_min_steps(4) != 0 and L.dc_get_split_k_min_steps() == 16


@pytest.mark.parametrize("co", [160, 192, 320])      # 320 >= 256 takes the 16-warp lean epilogue
def test_conv_residual_ragged_channel_tile(co, _gpu):
    # the second 128-channel tile is partly empty; residual prefetch registers of skipped chunks must
    # not leak into the next tile (several pixel tiles so every CTA sees both n-tiles)
    rng = np.random.default_rng(21)
    n, ci, h, w = 2, 64, 40, 37
    x = rng.standard_normal((n, ci, h, w)).astype(np.float32)
    wt = (rng.standard_normal((co, ci, 1, 1)) * np.sqrt(2.0 / ci)).astype(np.float32)
    a, b = _bn_params(rng, co)
    shortcut = rng.standard_normal((n, co, h, w)).astype(np.float32)
    ref = np.maximum(caffe_ref.convolution(x, wt, None, 1, 0, 1) * a.reshape(1, -1, 1, 1) + b.reshape(1, -1, 1, 1) + shortcut, 0)
    got = _gpu.conv_bn(x, wt, a, b, relu=True, residual_nchw=shortcut)
    assert np.abs(got - ref).max() < 1e-4


@pytest.mark.parametrize("n,ci,co,h,w,stride", [(2, 256, 128, 36, 64, 2), (1, 64, 512, 37, 51, 2), (2, 128, 64, 19, 23, 3), (1, 64, 128, 9, 300, 2)])
def test_strided_pointwise_conv_reads_through_tma_traversal_stride(n, ci, co, h, w, stride, _gpu):
    """The reference im2col's its strided 1x1 convs (res3a/res4a branch1 + branch2a, base_conv_layer.cpp:109-116); here
    the A tensor map traverses W and H with the stride.  Odd sizes: the last strided pixel is the image's last row/column."""
    rng = np.random.default_rng(n * 100 + ci + stride)
    x = np.maximum(rng.standard_normal((n, ci, h, w)), 0).astype(np.float32)
    wt = (rng.standard_normal((co, ci, 1, 1)) * 0.05).astype(np.float32)
    a = rng.uniform(0.5, 1.5, co).astype(np.float32)
    b = rng.normal(0, 0.1, co).astype(np.float32)
    got = _gpu.conv_bn(x, wt, a, b, relu=True, stride=stride)
    ref = np.maximum(caffe_ref.convolution(x, wt, None, stride, 0, 1) * a.reshape(1, -1, 1, 1) + b.reshape(1, -1, 1, 1), 0)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 1e-4


def test_strided_conv_rejects_unsupported_geometry(_gpu):
    L = libdc.lib()
    args = libdc.ConvArgs(x=1, n=1, h=8, w=8, cin=64, cout=64, kh=3, kw=3, pad=1, dilation=1, w_packed=1, scale=1, shift=1, residual=None,
                          relu=0, out_f32_rows=0, ldc=0, out=1, stride=2)
    assert L.dc_conv_forward(C.byref(args), None) != 0
    assert b"stride" in L.dc_last_error()


def test_conv_f32_rows_with_bias_heads(_gpu):
    # merged 1x1 heads: 512 -> 14+28+364 = 406 with bias, fp32 rows out (res3d_* layers)
    rng = np.random.default_rng(8)
    n, ci, co, h, w = 1, 512, 406, 9, 14
    x = rng.standard_normal((n, ci, h, w)).astype(np.float32)
    wt = (rng.standard_normal((co, ci, 1, 1)) * 0.01).astype(np.float32)
    bias = rng.normal(0, 0.1, co).astype(np.float32)
    ref = caffe_ref.convolution(x, wt, bias, 1, 0, 1)
    rows = _gpu.conv_bn(x, wt, np.ones(co, np.float32), bias, relu=False, f32_rows=True)
    got = rows[:, :co].reshape(n, h, w, co).transpose(0, 3, 1, 2)
    assert np.abs(got - ref).max() < 1e-5
    assert np.all(rows[:, co:] == 0)


def test_linearity_property_full_size(_gpu):
    # size-independent property at a BASELINE-size tile count: conv(x1 + x2) == conv(x1) + conv(x2)
    rng = np.random.default_rng(9)
    n, ci, co, h, w = 1, 512, 512, 45, 80          # res5 map of one 720p image
    wt = (rng.standard_normal((co, ci, 3, 3)) * np.sqrt(2.0 / (ci * 9))).astype(np.float32)
    x1 = rng.standard_normal((n, ci, h, w)).astype(np.float32)
    x2 = rng.standard_normal((n, ci, h, w)).astype(np.float32)
    one, zero = np.ones(co, np.float32), np.zeros(co, np.float32)
    y1 = _gpu.conv_bn(x1, wt, one, zero, pad=2, dil=2, relu=False)
    y2 = _gpu.conv_bn(x2, wt, one, zero, pad=2, dil=2, relu=False)
    y12 = _gpu.conv_bn(x1 + x2, wt, one, zero, pad=2, dil=2, relu=False)
    assert np.abs(y12 - (y1 + y2)).max() < 2e-4
    # and a spot check of 64 random outputs against a float64 dot product
    xp = np.pad(x1, ((0, 0), (0, 0), (2, 2), (2, 2)))
    for _ in range(64):
        c, yy, xx = rng.integers(co), rng.integers(h), rng.integers(w)
        patch = xp[0, :, yy:yy + 5:2, xx:xx + 5:2].astype(np.float64)
        assert abs((patch * wt[c]).sum() - y1[0, c, yy, xx]) < 1e-4


def test_conv1_stem(_gpu):
    L = libdc.lib()
    rng = np.random.default_rng(10)
    for (n, h, w) in ((1, 64, 64), (2, 75, 101)):
        x = dcutil.synth.images(n, h, w, seed=3)
        wt = (rng.standard_normal((64, 3, 7, 7)) * np.sqrt(2.0 / 147)).astype(np.float32)
        a, b = _bn_params(rng, 64)
        a = (a / 70).astype(np.float32)
        ref = np.maximum(caffe_ref.convolution(x, wt, None, 2, 3, 1) * a.reshape(1, -1, 1, 1) + b.reshape(1, -1, 1, 1), 0)
        wp = np.zeros((147, 64), np.float32)
        libdc.check(L.dc_pack_conv1_weight(dcutil.ptr(wt), dcutil.ptr(wp)))
        ho, wo = ref.shape[2:]
        out = torch.full((2, n, ho, wo, 64), float("nan"), dtype=torch.float16, device="cuda")
        dx, dw, da, db = _gpu.dev(x), _gpu.dev(wp), _gpu.dev(a), _gpu.dev(b)
        libdc.check(L.dc_conv1_forward(dx.data_ptr(), n, h, w, dw.data_ptr(), da.data_ptr(), db.data_ptr(),
                                       out.data_ptr(), _gpu.stream_ptr()))
        torch.cuda.synchronize()
        got = dcutil.np_join(out.cpu().numpy())
        assert np.abs(got - ref).max() < 1e-4, np.abs(got - ref).max()


def test_conv1_stem_tensor_core(_gpu):
    """dc_conv1_tc_forward: space-to-depth + overlapping-window tensor map + tcgen05 conv, vs the oracle."""
    L = libdc.lib()
    rng = np.random.default_rng(14)
    for (n, h, w) in ((1, 64, 64), (2, 75, 101), (3, 150, 230), (1, 720, 1280)):      # the last two: enough tiles for TALL mode
        x = dcutil.synth.images(n, h, w, seed=3)
        wt = (rng.standard_normal((64, 3, 7, 7)) * np.sqrt(2.0 / 147)).astype(np.float32)
        a, b = _bn_params(rng, 64)
        a = (a / 70).astype(np.float32)
        packed = np.zeros((2, 64, 256), np.uint16)
        rs = np.zeros(64, np.float32)
        libdc.check(L.dc_pack_conv1_tc_weight(dcutil.ptr(wt), dcutil.ptr(packed), dcutil.ptr(rs)))
        ho, wo = (h + 1) // 2, (w + 1) // 2
        ws = torch.empty(L.dc_conv1_tc_workspace_bytes(n, h, w), dtype=torch.uint8, device="cuda")
        out = torch.full((2, n, ho, wo, 64), float("nan"), dtype=torch.float16, device="cuda")
        dx, dw, da, db = _gpu.dev(x), _gpu.dev(packed), _gpu.dev(a * rs), _gpu.dev(b)
        libdc.check(L.dc_conv1_tc_forward(dx.data_ptr(), n, h, w, dw.data_ptr(), da.data_ptr(), db.data_ptr(), ws.data_ptr(),
                                          out.data_ptr(), _gpu.stream_ptr()))
        torch.cuda.synchronize()
        got = dcutil.np_join(out.cpu().numpy())
        if h * w > 100000:      # full 720p: fp64 spot checks instead of the whole oracle conv
            xp = np.pad(x[0], ((0, 0), (3, 3), (3, 3))).astype(np.float64)
            for _ in range(200):
                c, oy, ox = rng.integers(64), rng.integers(ho), rng.integers(wo)
                ref = max((xp[:, 2 * oy:2 * oy + 7, 2 * ox:2 * ox + 7] * wt[c]).sum() * a[c] + b[c], 0)
                assert abs(got[0, c, oy, ox] - ref) < 1e-4
        else:
            ref = np.maximum(caffe_ref.convolution(x, wt, None, 2, 3, 1) * a.reshape(1, -1, 1, 1) + b.reshape(1, -1, 1, 1), 0)
            assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-4, np.abs(got - ref).max()


def test_maxpool_bit_exact(_gpu):
    L = libdc.lib()
    rng = np.random.default_rng(11)
    for (n, c, h, w) in ((1, 64, 32, 32), (2, 64, 37, 50), (1, 8, 5, 3)):
        x = rng.standard_normal((n, c, h, w)).astype(np.float32)
        xs = dcutil.np_split(x)
        xj = dcutil.np_join(xs)                      # the values the kernel actually sees
        ref = caffe_ref.max_pool(xj, 3, 2)
        ho, wo = ref.shape[2:]
        assert (ho, wo) == (L.dc_pool_out_size(h, 3, 2), L.dc_pool_out_size(w, 3, 2))
        out = torch.zeros((2, n, ho, wo, c), dtype=torch.float16, device="cuda")
        dx = _gpu.dev(xs)
        libdc.check(L.dc_maxpool_forward(dx.data_ptr(), n, h, w, c, 3, 2, out.data_ptr(), _gpu.stream_ptr()))
        torch.cuda.synchronize()
        assert np.array_equal(dcutil.np_join(out.cpu().numpy()), ref)


def test_subsample_and_layout_roundtrip(_gpu):
    L = libdc.lib()
    rng = np.random.default_rng(12)
    n, c, h, w = 2, 256, 13, 18
    x = rng.standard_normal((n, c, h, w)).astype(np.float32)
    dx = _gpu.dev(x)
    sp = torch.zeros((2, n, h, w, c), dtype=torch.float16, device="cuda")
    libdc.check(L.dc_nchw_to_split(dx.data_ptr(), n, c, h, w, sp.data_ptr(), _gpu.stream_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(sp.cpu().numpy().view(np.uint16), dcutil.np_split(x).view(np.uint16))
    back = torch.zeros((n, c, h, w), dtype=torch.float32, device="cuda")
    libdc.check(L.dc_split_to_nchw(sp.data_ptr(), n, c, h, w, back.data_ptr(), _gpu.stream_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(back.cpu().numpy(), dcutil.np_join(dcutil.np_split(x)))
    ho, wo = (h + 1) // 2, (w + 1) // 2
    sub = torch.zeros((2, n, ho, wo, c), dtype=torch.float16, device="cuda")
    libdc.check(L.dc_subsample_forward(sp.data_ptr(), n, h, w, c, 2, sub.data_ptr(), _gpu.stream_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(sub.cpu().numpy(), sp.cpu().numpy()[:, :, ::2, ::2, :])


def _gemm_channel_major(_gpu, x_nchw, packed, scale, shift, cout):
    """dc_conv_forward with out_f32_rows = 2 (operands swapped): returns fp32 [rows][ld]."""
    L = libdc.lib()
    n, ci, h, w = x_nchw.shape
    rows = packed.shape[1]
    ld = (n * h * w + 31) // 32 * 32
    xs = _gpu.dev(dcutil.np_split(x_nchw))
    out = torch.full((rows, ld), float("nan"), dtype=torch.float32, device="cuda")
    dwp, dsc, dsh = _gpu.dev(packed), _gpu.dev(scale), _gpu.dev(shift)
    args = libdc.ConvArgs(x=xs.data_ptr(), n=n, h=h, w=w, cin=ci, cout=cout, kh=1, kw=1, pad=0, dilation=1,
                          w_packed=dwp.data_ptr(), scale=dsc.data_ptr(), shift=dsh.data_ptr(), residual=None,
                          relu=0, out_f32_rows=2, ldc=ld, out=out.data_ptr())
    libdc.check(L.dc_conv_forward(C.byref(args), _gpu.stream_ptr()))
    torch.cuda.synchronize()
    return out, ld


def test_deconv_head_pipeline(_gpu):
    """Deconvolution 3x3/2 + Crop + Eltwise(+1x1 skip conv with bias) + Sigmoid, as three ABI calls
    (two channel-major GEMMs + finish), vs the oracle's layer-by-layer result (deconv_layer.cpp,
    crop_layer.cpp, eltwise_layer.cpp, sigmoid_layer.cpp)."""
    L = libdc.lib()
    rng = np.random.default_rng(13)
    # n*h*w not a multiple of 128: ragged pixel tiles; (dh, dw): skip map = (2h+dh) x (2w+dw) (the reference's Crop wants
    # it strictly smaller than the (2h+1) x (2w+1) deconv output) -- even widths take the float2 path, odd ones the scalar
    for (n, h, w, dh, dw) in ((2, 6, 7, 0, 0), (1, 11, 13, 0, 0), (2, 5, 9, 0, -1), (1, 7, 4, -1, 0), (1, 3, 3, -1, -1), (1, 4, 6, -2, -3)):
        c5, c3 = 256, 128
        ho, wo = 2 * h + dh, 2 * w + dw
        heads = (("pose", 14, True), ("locref", 28, False), ("next", 100, False))
        x5 = np.maximum(rng.standard_normal((n, c5, h, w)), 0).astype(np.float32)
        x3 = np.maximum(rng.standard_normal((n, c3, ho, wo)), 0).astype(np.float32)
        wd = [(rng.standard_normal((c5, co, 3, 3)) * 0.05).astype(np.float32) for _, co, _ in heads]
        bd = [rng.normal(0, 0.1, co).astype(np.float32) for _, co, _ in heads]
        ws = [(rng.standard_normal((co, c3, 1, 1)) * 0.05).astype(np.float32) for _, co, _ in heads]
        bs = [rng.normal(0, 0.1, co).astype(np.float32) for _, co, _ in heads]
        ctot = sum(co for _, co, _ in heads)
        packed, rs = dcutil.pack_deconv(np.concatenate(wd, axis=1))
        col, ldcol = _gemm_channel_major(_gpu, x5, packed, rs, np.zeros_like(rs), ctot * 9)
        packed2, rs2 = dcutil.pack_conv(np.concatenate(ws, axis=0))
        shift2 = np.zeros_like(rs2)
        shift2[:ctot] = np.concatenate(bs) + np.concatenate(bd)
        skip, ldskip = _gemm_channel_major(_gpu, x3, packed2, rs2, shift2, ctot)
        # the channel-major GEMM result is Caffe's col buffer: W^T x per pixel
        ref_col = np.einsum("nchw,cr->rnhw", x5.astype(np.float64), np.concatenate(wd, axis=1).reshape(c5, -1).astype(np.float64))
        assert np.abs(col.cpu().numpy()[:ctot * 9, :n * h * w] - ref_col.reshape(ctot * 9, -1)).max() < 1e-4
        off = 0
        for i, (name, co, sig) in enumerate(heads):
            up = caffe_ref.deconvolution(x5, wd[i], bd[i], 2, 0, 1)
            sk = caffe_ref.convolution(x3, ws[i], bs[i], 1, 0, 1)
            ref = caffe_ref.eltwise_sum([sk, caffe_ref.crop(up, sk)])
            if sig:
                ref = caffe_ref.sigmoid(ref)
            out = torch.full((n, co, ho, wo), float("nan"), dtype=torch.float32, device="cuda")
            libdc.check(L.dc_head_finish(col.data_ptr(), ldcol, off * 9, skip.data_ptr(), ldskip, off, out.data_ptr(), n, co, h, w,
                                         ho, wo, int(sig), _gpu.stream_ptr()))
            torch.cuda.synchronize()
            assert np.abs(out.cpu().numpy() - ref).max() < 2e-5, name
            off += co


def test_pose_from_maps_matches_demo_readout(_gpu):
    L = libdc.lib()
    rng = np.random.default_rng(15)
    for (n, h, w, scale) in ((2, 90, 160, 1.0), (1, 33, 47, 0.5), (3, 8, 8, 1.5)):
        prob = rng.random((n, 14, h, w)).astype(np.float32)
        prob[0, 3] = 0.25                      # a constant map: arg-max must be index 0 like np.argmax
        prob[0, 5, h // 2, w // 3] = prob[0, 5, h - 1, w - 1] = 2.0       # tie: the first one wins
        loc = rng.standard_normal((n, 28, h, w)).astype(np.float32)
        out = torch.zeros((n, 5, 14), dtype=torch.float32, device="cuda")
        dp, dl = _gpu.dev(prob), _gpu.dev(loc)
        libdc.check(L.dc_pose_from_maps(dp.data_ptr(), dl.data_ptr(), n, 14, h, w, 8.0, float(np.sqrt(53.0)), scale, out.data_ptr(),
                                        _gpu.stream_ptr()))
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        for i in range(n):
            ref = caffe_ref.pose_from_mats(prob[i], loc[i], scale=scale)
            assert np.abs(got[i] - ref).max() < 1e-3, (i, np.abs(got[i] - ref).max())
            assert np.array_equal(got[i][2], ref[2].astype(np.float32))      # confidences are exact copies


def test_launch_counter_moves(_gpu):
    before = libdc.lib().dc_launch_count()
    test_subsample_and_layout_roundtrip(_gpu)
    assert libdc.lib().dc_launch_count() >= before + 3
