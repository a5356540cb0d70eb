"""Behaviour either side of the fused plan that callers of the Caffe API can observe (GPU): which blobs a forward leaves
current, partial forwards, head groups trimmed by the prototxt, the debug_info probes and the `caffe time` tool."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import dcutil
import netutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def test_intermediate_blobs_are_never_served_stale(tmp_path):
    """The reference fills every blob on every forward (net.cpp:565-581); the fused plan only the outputs.  Reading an
    intermediate after a fused forward raises instead of returning old data; asking for it (pycaffe's forward(blobs=[...]))
    runs that call layer by layer and returns the reference's value; a partial forward from the middle is refused until the
    bottoms it needs have been materialised."""
    caffe = dcutil.caffe_module()
    path, weights = netutil.build(tmp_path, (1, 1, 1, 1), 64, 64)
    x = dcutil.synth.images(1, 64, 64, seed=21)
    ref = netutil.oracle_forward(path, weights, x, want={"res3a_relu", "pool1"})
    net = netutil.product_net(path, weights)
    out = netutil.product_forward(net, x)
    assert net.fused_last_forward
    with pytest.raises(caffe._caffe.CaffeError, match="not written by the last forward"):
        net.blobs["res3a"].data
    with pytest.raises(caffe._caffe.CaffeError, match="did not materialise"):
        net.forward(start="res4a_branch1")
    got = net.forward(blobs=["res3a", "pool1"])
    assert not net.fused_last_forward
    assert netutil.max_err(np.array(got["res3a"]), ref["res3a_relu"]) < 1e-4
    assert netutil.max_err(np.array(got["pool1"]), ref["pool1"]) < 1e-4
    assert netutil.max_err(np.array(got["prob"]), out["prob"]) < 1e-4
    net.forward(start="res4a_branch1")                       # now every bottom is current
    net.blobs["res3a"].data                                    # and so is this
    again = netutil.product_forward(net, x)                    # fused again
    assert net.fused_last_forward
    for k in out:
        assert np.array_equal(again[k], out[k]), k


def test_net_trimmed_to_part_and_locref_heads_stays_fused(tmp_path):
    """A deploy net without the 364-channel next_pred head (the demo never reads it, estimate_pose.py:231): 14 + 28 = 42 merged
    head outputs, fewer than one 128-row GEMM tile -- the planner pads instead of refusing."""
    path = dcutil.write_prototxt(tmp_path, stages=(1, 1, 1, 1), height=64, width=64, heads=(("pose", 14), ("locref", 28)))
    full_path, weights = netutil.build(tmp_path, (1, 1, 1, 1), 64, 64)
    w2 = {k: v for k, v in weights.items() if "next" not in k}
    x = dcutil.synth.images(2, 64, 64, seed=22)
    ref = netutil.oracle_forward(full_path, weights, x)
    net = netutil.product_net(path, w2)
    got = netutil.product_forward(net, x)
    assert net.fused_last_forward, net.fusion_diagnostic
    assert sorted(got) == ["loc_pred", "prob"]
    for k in got:
        assert netutil.max_err(got[k], ref[k]) < 1e-4, k


def test_skipped_outputs_leave_the_other_heads_bitwise_unchanged(tmp_path):
    """net.skip_outputs(['next_pred']): the full deploy net, but the 364-channel head is neither computed nor written; prob and
    loc_pred keep their per-element K chains (same weight rows, other row tiles), so they are bitwise what the full plan gives.
    Reading the skipped blob raises instead of handing out stale data; clearing the list restores the full plan."""
    path, weights = netutil.build(tmp_path, (1, 2, 2, 1), 96, 80)
    x = dcutil.synth.images(3, 96, 80, seed=5)
    net = netutil.product_net(path, weights)
    full = netutil.product_forward(net, x)
    launches_full = net.last_forward_launches
    net.skip_outputs(["next_pred"])
    got = netutil.product_forward(net, x)
    assert net.fused_last_forward, net.fusion_diagnostic
    assert sorted(got) == ["loc_pred", "prob"]
    assert net.last_forward_launches == launches_full - 1            # one HeadFinish less; the two GEMMs shrink
    for k in got:
        assert np.array_equal(got[k], full[k]), (k, netutil.max_err(got[k], full[k]))
    with pytest.raises(dcutil.caffe_module()._caffe.CaffeError, match="not written by the last forward"):
        net.blobs["next_pred"].data
    net.skip_outputs([])
    again = netutil.product_forward(net, x)
    assert sorted(again) == ["loc_pred", "next_pred", "prob"]
    for k in again:
        assert np.array_equal(again[k], full[k]), k


def test_debug_info_probes_match_the_oracle_blob_by_blob(tmp_path):
    """debug_info (net.cpp:648-735): mean|x| of every top after every layer, here compared with the same statistic of the CPU
    oracle's blobs -- a per-blob check of the whole per-layer plugin path."""
    path, weights = netutil.build(tmp_path, (1, 1, 1, 1), 64, 64)
    x = dcutil.synth.images(1, 64, 64, seed=23)
    net = netutil.product_net(path, weights)
    net.set_debug_info(True)
    netutil.product_forward(net, x)
    assert not net.fused_last_forward                          # debug_info runs layer by layer, like the reference
    probes = net.debug_info()
    names = [p[0] for p in probes]
    ref = netutil.oracle_forward(path, weights, x, want=set(names))
    assert len(probes) >= 60
    checked = 0
    for layer, blob, mean_abs in probes:
        if layer not in ref:
            continue
        want = float(np.abs(ref[layer]).mean())
        assert abs(mean_abs - want) <= 1e-4 * max(1.0, want), (layer, blob, mean_abs, want)
        checked += 1
    assert checked >= 60
    net.set_debug_info(False)
    netutil.product_forward(net, x)
    assert net.fused_last_forward


def test_caffe_time_tool_reports_every_fused_step(tmp_path):
    """tools/caffe_time.py = `caffe time` (tools/caffe.cpp:302-388) over the fused plan."""
    path = dcutil.write_prototxt(tmp_path, stages=(1, 1, 1, 1), height=64, width=64)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "caffe_time.py"), "--model", path, "--batch", "2", "--iterations", "3"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "*** Benchmark begins ***" in r.stdout and "*** Benchmark ends ***" in r.stdout
    steps = re.findall(r"^(\w+)\s+(\S+)\s+forward:\s+([0-9.]+) ms", r.stdout, re.M)
    kinds = {s[0] for s in steps}
    assert {"Conv1", "ConvBN", "MaxPool", "HeadGemm", "HeadFinish"} <= kinds, r.stdout[-3000:]
    total = float(re.search(r"Average Forward pass: ([0-9.]+) ms", r.stdout).group(1))
    assert abs(total - sum(float(s[2]) for s in steps)) < 0.05 * total + 0.01
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "caffe_time.py"), "--model", path, "--iterations", "2", "--per-layer"],
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r2.returncode == 0 and "per-layer path" in r2.stdout, r2.stdout[-2000:]
