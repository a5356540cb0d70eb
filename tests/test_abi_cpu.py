"""CPU-side checks of the C ABI: the library loads without a GPU, exports every symbol the
header declares, and the load-time weight transforms are bit-exact against numpy."""
import os
import re

import numpy as np

import dcutil
from dcutil import libdc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "deepcut_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(dc_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    L = libdc.lib()
    for name in declared:
        assert hasattr(L, name), "symbol %s declared in the header but not exported" % name
    assert set(declared) == set(libdc.exported_symbols()), "ctypes binding and header disagree"
    assert L.dc_version() >= 100
    assert L.dc_device_count() >= 0


def test_compute_entry_points_fail_loudly_without_gpu():
    L = libdc.lib()
    if L.dc_device_count() > 0:
        return
    rc = L.dc_init(0)
    assert rc == 3, "dc_init must report DC_ERR_NO_DEVICE on a CPU-only host"
    assert b"no CPU path" in L.dc_last_error()
    a = libdc.ConvArgs()
    assert L.dc_conv_forward(a, None) != 0


def test_split_k_policy_setters_validate_their_arguments():
    # host-side state only: no GPU needed
    L = libdc.lib()
    assert L.dc_get_split_k() == 4 and L.dc_get_split_k_min_steps() == 16          # defaults (env DC_SPLIT_K / DC_SPLIT_K_MIN_STEPS unset)
    for bad in (0, 3, 8, -1):
        assert L.dc_set_split_k(bad) != 0 and b"dc_set_split_k" in L.dc_last_error()
    assert L.dc_set_split_k_min_steps(7) != 0
    assert L.dc_set_split_k(2) == 0 and L.dc_get_split_k() == 2
    assert L.dc_set_split_k_min_steps(36) == 0 and L.dc_get_split_k_min_steps() == 36
    assert L.dc_set_split_k(4) == 0 and L.dc_set_split_k_min_steps(16) == 0
    # one 128 x 128 fp32 tile per SM bounds any split launch's scratch
    assert L.dc_splitk_workspace_bytes() >= 148 * 128 * 128 * 4
    a = libdc.ConvArgs()
    assert a.splitk_workspace is None and a.splitk_workspace_bytes == 0           # a zeroed struct never splits


def test_pack_conv_weight_matches_numpy_split():
    rng = np.random.default_rng(0)
    for (co, ci, k) in ((64, 64, 1), (128, 64, 3), (14, 512, 1), (300, 64, 3)):
        w = (rng.standard_normal((co, ci, k, k)) * rng.choice([1e-3, 0.05, 3.0])).astype(np.float32)
        w[co // 2] *= 1e-4     # a tiny row: the power-of-two row scale must rescue its lo plane
        packed, rs = dcutil.pack_conv(w)
        rows = libdc.lib().dc_packed_rows(co)
        assert packed.shape == (2, rows, k * k * ci) and rows % libdc.lib().dc_tile_n(co) == 0
        # K order is (tap, ci); rows scaled by an exact power of two
        wk = w.transpose(0, 2, 3, 1).reshape(co, -1)
        s = 1.0 / rs[:co]
        assert np.all(np.log2(s) == np.round(np.log2(s)))
        mx = np.abs(wk * s[:, None]).max(axis=1)
        assert np.all((mx >= 512) & (mx < 1024))
        v = wk * s[:, None].astype(np.float32)
        hi = v.astype(np.float16)
        lo = (v - hi.astype(np.float32)).astype(np.float16)
        assert np.array_equal(packed[0, :co].view(np.float16).view(np.uint16), hi.view(np.uint16))
        assert np.array_equal(packed[1, :co].view(np.float16).view(np.uint16), lo.view(np.uint16))
        assert not packed[:, co:].any() and np.all(rs[co:] == 1.0)
        rec = (hi.astype(np.float64) + lo.astype(np.float64)) * rs[:co, None]
        assert np.abs(rec - wk).max() <= 2.0 ** -21 * np.abs(wk).max(axis=1).max()


def test_pack_deconv_weight_layout():
    rng = np.random.default_rng(1)
    ci, co = 128, 14
    w = rng.standard_normal((ci, co, 3, 3)).astype(np.float32) * 0.01
    packed, rs = dcutil.pack_deconv(w)
    rows = libdc.lib().dc_packed_rows(co * 9)
    assert packed.shape == (2, rows, ci)
    wk = w.reshape(ci, co * 9).T          # row = co*9 + p*3 + q, K = ci
    rec = (packed[0].view(np.float16).astype(np.float64) + packed[1].view(np.float16).astype(np.float64)) * rs[:, None]
    assert np.abs(rec[:co * 9] - wk).max() < 1e-8
    assert not rec[co * 9:].any()


def test_fold_bn_scale_matches_reference_formula():
    # BatchNorm inference + Scale as the reference computes them (batch_norm_layer.cpp:86-149,
    # scale_layer.cpp:120-133) vs the folded a*x+b: equal to a few ulp.
    from oracle import caffe_ref
    rng = np.random.default_rng(2)
    c = 96
    bn = [rng.normal(0, 0.3, c).astype(np.float32), rng.uniform(0.2, 3, c).astype(np.float32), np.array([2.5], np.float32)]
    sc = [rng.uniform(0.5, 1.5, c).astype(np.float32), rng.normal(0, 0.2, c).astype(np.float32)]
    a, b = dcutil.fold_bn(bn, sc)
    x = rng.standard_normal((2, c, 5, 7)).astype(np.float32) * 3
    ref = caffe_ref.scale_bias(caffe_ref.batch_norm_global(x, bn[0], bn[1], bn[2][0]), sc[0], sc[1])
    got = x * a.reshape(1, -1, 1, 1) + b.reshape(1, -1, 1, 1)
    assert np.abs(got - ref).max() < 5e-6
    # scale_factor == 0 -> statistics are zeroed (batch_norm_layer.cpp:88-89)
    a0, b0 = dcutil.fold_bn([bn[0], bn[1], np.array([0.0], np.float32)], None)
    assert np.allclose(a0, 1.0 / np.sqrt(1e-5), rtol=1e-6) and np.all(b0 == 0)


def test_conv1_pack_and_pool_size():
    L = libdc.lib()
    w = np.arange(64 * 147, dtype=np.float32).reshape(64, 3, 7, 7)
    out = np.zeros((147, 64), np.float32)
    libdc.check(L.dc_pack_conv1_weight(dcutil.ptr(w), dcutil.ptr(out)))
    assert np.array_equal(out, w.reshape(64, 147).T)
    from oracle import caffe_ref
    for size in (5, 6, 7, 128, 129, 344, 360, 640):
        assert L.dc_pool_out_size(size, 3, 2) == caffe_ref.pool_out_size(size, 3, 0, 2)


def test_conv1_tc_weight_mapping_reproduces_7x7_stride2():
    """dc_pack_conv1_tc_weight lays the 7x7/2 filter out as a 4x4 stride-1 filter over the 2x2
    space-to-depth image (16 ch per pixel).  Emulate stem_s2d_kernel + that convolution in numpy and
    compare with the oracle's 7x7/2 convolution (conv_layer.cpp:8-40)."""
    from oracle import caffe_ref
    L = libdc.lib()
    rng = np.random.default_rng(5)
    w = (rng.standard_normal((64, 3, 7, 7)) * 0.1).astype(np.float32)
    packed = np.zeros((2, 64, 256), np.uint16)
    rs = np.zeros(64, np.float32)
    libdc.check(L.dc_pack_conv1_tc_weight(dcutil.ptr(w), dcutil.ptr(packed), dcutil.ptr(rs)))
    wk = (packed[0].view(np.float16).astype(np.float64) + packed[1].view(np.float16).astype(np.float64)) * rs[:, None]   # [64][256]
    assert abs(np.abs(wk).sum() - np.abs(w.astype(np.float64)).sum()) < 1e-3          # every tap placed exactly once
    for (h, wd) in ((12, 16), (11, 13)):
        x = rng.standard_normal((1, 3, h, wd)).astype(np.float32)
        h2, w2 = (h + 1) // 2, (wd + 1) // 2
        s2d = np.zeros((h2, w2 + 3, 16))
        for py in range(2):
            for px in range(2):
                for ci in range(3):
                    sub = x[0, ci, py::2, px::2]
                    s2d[:sub.shape[0], 2:2 + sub.shape[1], (py * 2 + px) * 3 + ci] = sub
        out = np.zeros((64, h2, w2))
        for oy in range(h2):
            for ox in range(w2):
                k = np.zeros(256)
                for p in range(4):
                    yy = oy + p - 2
                    if 0 <= yy < h2:
                        k[p * 64:(p + 1) * 64] = s2d[yy, ox:ox + 4, :].reshape(64)
                out[:, oy, ox] = wk @ k
        ref = caffe_ref.convolution(x, w, None, 2, 3, 1)[0]
        assert out.shape == ref.shape and np.abs(out - ref).max() < 1e-4
