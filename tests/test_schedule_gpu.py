"""The L2-resident chunked schedule on the GPU (dc_engine.cpp PlanSchedule): running a stage's blocks sub-batch by sub-batch,
with segment-local tensors holding one pass only and block outputs written in place over their shortcut, must not change a
single bit of the outputs -- every output element keeps its own K chain whatever the launch covers."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import dcutil
import gpuharness
import netutil


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    gpuharness.init()
    L = dcutil.libdc.lib()
    before = L.dc_get_split_k()
    dcutil.libdc.check(L.dc_set_split_k(1))        # split-K decisions depend on the launch geometry; off = bitwise comparable
    yield
    dcutil.libdc.check(L.dc_set_split_k(before))


@pytest.mark.parametrize("case", [
    # (n, ci, co, h, w, k, pad, dil, i0, cn)
    (5, 64, 64, 20, 24, 3, 1, 1, 1, 2),          # 3x3: per-image tile grids
    (4, 128, 128, 9, 11, 3, 2, 2, 3, 1),         # dilated, last image only
    (5, 256, 64, 13, 10, 1, 0, 1, 2, 3),         # flat 1x1: 130 pixels per image, sub-batch starts/ends inside 128-pixel tiles
    (3, 64, 256, 8, 8, 1, 0, 1, 0, 2),
])
def test_sub_batch_launch_equals_full_launch(case):
    n, ci, co, h, w, k, pad, dil, i0, cn = case
    rng = np.random.default_rng(sum(case))
    x = np.maximum(rng.standard_normal((n, ci, h, w)), 0).astype(np.float32)
    wt = (rng.standard_normal((co, ci, k, k)) * (2.0 / (ci * k * k)) ** 0.5).astype(np.float32)
    a = rng.uniform(0.5, 1.5, co).astype(np.float32)
    b = rng.normal(0, 0.1, co).astype(np.float32)
    full = gpuharness.conv_bn(x, wt, a, b, pad=pad, dil=dil, relu=True, split_k_workspace=False)       # fp32 NCHW = hi + lo
    part = gpuharness.conv_bn_subbatch(x, wt, a, b, i0, cn, pad=pad, dil=dil, relu=True)
    sel = np.zeros(n, bool)
    sel[i0:i0 + cn] = True
    assert np.array_equal(dcutil.np_join(part[:, sel]), full[sel])
    assert np.isnan(part[:, ~sel].astype(np.float32)).all(), "the launch wrote outside its sub-batch"


@pytest.mark.parametrize("co,k,h,w", [(256, 1, 48, 64), (256, 1, 12, 20), (128, 3, 12, 20)])
def test_block_output_in_place_over_the_shortcut(co, k, h, w):
    # out == residual: every element is read, then written, by the same epilogue warp.  48x64: enough tiles for the CTA-pair
    # kernel with the lean 16-warp epilogue (residual through cp.async); the small cases run the 8-warp epilogue
    n, ci = 4, 64
    rng = np.random.default_rng(co + k)
    x = np.maximum(rng.standard_normal((n, ci, h, w)), 0).astype(np.float32)
    r = rng.standard_normal((n, co, h, w)).astype(np.float32)
    wt = (rng.standard_normal((co, ci, k, k)) * (2.0 / (ci * k * k)) ** 0.5).astype(np.float32)
    a = rng.uniform(0.5, 1.5, co).astype(np.float32)
    b = rng.normal(0, 0.1, co).astype(np.float32)
    want = gpuharness.conv_bn(x, wt, a, b, pad=k // 2, relu=True, residual_nchw=r, split_k_workspace=False)
    got = gpuharness.conv_bn_subbatch(x, wt, a, b, 1, 2, pad=k // 2, relu=True, residual_nchw=r, inplace=True)
    rs = dcutil.np_split(r)
    assert np.array_equal(dcutil.np_join(got[:, 1:3]), want[1:3])
    for i in (0, 3):      # images outside the sub-batch still hold the shortcut
        assert np.array_equal(got[:, i].view(np.uint16), rs[:, i].view(np.uint16))


def _forward(tmp_path, monkeypatch, stages, x, plan, inplace):
    monkeypatch.setenv("DC_CHUNK_PLAN", plan)
    monkeypatch.setenv("DC_INPLACE_RESIDUAL", inplace)
    path, weights = netutil.build(tmp_path, stages, x.shape[2], x.shape[3])
    net = netutil.product_net(path, weights)
    out = netutil.product_forward(net, x)
    assert net.fused_last_forward, net.fusion_diagnostic
    again = netutil.product_forward(net, x)          # CUDA-graph replay of the same schedule
    for k in out:
        assert np.array_equal(out[k], again[k]), k
    return path, weights, out, net.last_forward_launches


def test_chunked_schedule_is_bitwise_identical_and_matches_the_reference(tmp_path, monkeypatch):
    stages = (2, 2, 3, 2)
    x = dcutil.synth.images(5, 48, 80, seed=3)
    path, weights, base, l0 = _forward(tmp_path, monkeypatch, stages, x, "0,0,0,0", "0")
    ref = netutil.reference_forward(path, weights, x, want=["prob", "loc_pred", "next_pred"]) if netutil.reference_available() \
        else netutil.oracle_forward(path, weights, x)
    for k in ("prob", "loc_pred", "next_pred"):
        assert netutil.max_err(base[k], ref[k]) < 1e-4, k
    for plan, inplace in (("0,0,0,0", "1"), ("2,2,2,2", "1"), ("1,3,2,4", "1"), ("2,1,2,1", "0")):
        _, _, got, l1 = _forward(tmp_path, monkeypatch, stages, x, plan, inplace)
        for k in ("prob", "loc_pred", "next_pred"):
            assert np.array_equal(got[k], base[k]), (plan, inplace, k)
        if plan != "0,0,0,0":
            assert l1 > l0          # the segments really ran in several passes


def test_default_budget_chunks_a_batch_that_overflows_l2(tmp_path, monkeypatch):
    # a 2 MiB budget makes the default policy chunk even this small batch; results as without chunking
    stages = (1, 2, 2, 1)
    x = dcutil.synth.images(6, 64, 64, seed=5)
    monkeypatch.setenv("DC_L2_CHUNK_MB", "0")
    _, _, base, l0 = _forward(tmp_path, monkeypatch, stages, x, "", "1")
    monkeypatch.setenv("DC_L2_CHUNK_MB", "0.5")
    monkeypatch.delenv("DC_CHUNK_PLAN")
    path, weights = netutil.build(tmp_path, stages, 64, 64)
    net = netutil.product_net(path, weights)
    got = netutil.product_forward(net, x)
    assert net.last_forward_launches > l0, net.describe_plan()
    for k in ("prob", "loc_pred", "next_pred"):
        assert np.array_equal(got[k], base[k]), k
